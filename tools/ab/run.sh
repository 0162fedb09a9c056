set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_handoff.py tests/test_conversion.py -x -q 2>&1 | tail -15
for c in 1 0; do JP_MOVE_TIMING=1 timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify $c > gpurun_out/tp_cls$c.log 2>&1; tail -7 gpurun_out/tp_cls$c.log | grep -v jp_move; grep "jp_move" gpurun_out/tp_cls$c.log | tail -1; done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
