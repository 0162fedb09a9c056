python -m pytest tests -m gpu -x -q 2>&1 | tail -2
export JP_MOVE_TIMING=1
python tools/time_phases.py --cells 128 --steps 8 2>&1 | grep -E "^lib|^move|^p2g|^phase|jp_move" | tail -5
python tools/time_phases.py --cells 256 --steps 8 2>&1 | grep -E "^lib|^adv|^move|^p2g|^phase|jp_move" | tail -6
