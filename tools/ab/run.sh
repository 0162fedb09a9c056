set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_handoff.py -x -q 2>&1 | tail -5
for c in 1 0; do JP_MOVE_TIMING=1 timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify $c > gpurun_out/tp_cls$c.log 2>&1; tail -7 gpurun_out/tp_cls$c.log | grep -E "advect|move|classify"; grep "jp_move" gpurun_out/tp_cls$c.log | tail -1; done
for v in B C D; do JUSTPIC_LIB=tools/ab/var$v.so JP_MOVE_TIMING=1 timeout 300 python tools/time_phases.py --cells 256 --steps 8 > gpurun_out/tp_var$v.log 2>&1; echo "variant $v"; grep "jp_move" gpurun_out/tp_var$v.log | tail -1; grep checksum gpurun_out/tp_var$v.log; done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
