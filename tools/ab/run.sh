python -m pytest tests -m gpu -x -q 2>&1 | tail -2
export JP_MOVE_TIMING=1
python tools/time_phases.py --cells 256 --steps 6 2>&1 | grep -E "^lib|^move|jp_move" | tail -3
JUSTPIC_LIB=tools/ab/lib_C3.so python tools/time_phases.py --cells 256 --steps 6 2>&1 | grep -E "^lib|^move|jp_move" | tail -3
JP_MOVE_CLASSIFY2=1 JUSTPIC_LIB=tools/ab/lib_C3.so python tools/time_phases.py --cells 256 --steps 6 2>&1 | grep -E "^lib|^move|jp_move" | tail -3
