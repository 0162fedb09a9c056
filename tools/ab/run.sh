export JP_MOVE_TIMING=1
for L in lib_P1 lib_P2 lib_P3; do echo $L; JUSTPIC_LIB=tools/ab/$L.so python tools/time_phases.py --cells 256 --steps 5 2>&1 | grep -E "jp_move|checksum" | tail -2; done
python tools/time_phases.py --cells 256 --steps 5 2>&1 | grep -E "jp_move|checksum" | tail -2
