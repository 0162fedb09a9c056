python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/time_phases.py --cells 128 --steps 10 2>&1 | grep -E "^lib|^advect|checksum"
for L in lib_W1 lib_W2; do JUSTPIC_LIB=tools/ab/$L.so python tools/time_phases.py --cells 128 --steps 10 2>&1 | grep -E "^lib|^advect|checksum"; done
python tools/time_phases.py --cells 256 --steps 8 2>&1 | grep -E "^lib|^advect|^move|^p2g|^phase|checksum"
