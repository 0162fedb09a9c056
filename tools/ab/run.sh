set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_handoff.py -x -q 2>&1 | tail -15
for c in 0 1; do JP_MOVE_TIMING=1 timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify $c > gpurun_out/tp_cls$c.log 2>&1; tail -8 gpurun_out/tp_cls$c.log; grep "jp_move" gpurun_out/tp_cls$c.log | tail -2; done
JUSTPIC_LIB=tools/ab/old.so JP_MOVE_TIMING=1 timeout 300 python tools/time_phases.py --cells 256 --steps 8 > gpurun_out/tp_old.log 2>&1; tail -8 gpurun_out/tp_old.log; grep "jp_move" gpurun_out/tp_old.log | tail -2
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
