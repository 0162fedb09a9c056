python -m pytest tests -m gpu -x -q -k "policy or trajectory" 2>&1 | tail -2
for P in reference compact; do python tools/time_phases.py --cells 128 --steps 16 --policy $P 2>&1 | grep -E "^policy|^advect|^move|^p2g|^phase"; done
