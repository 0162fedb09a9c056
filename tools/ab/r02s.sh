set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_advection_interpolants.py tests/test_gpu_parity.py -m gpu -x -q -k "interpolants or wide" 2>&1 | tail -6
( timeout 300 python tools/time_interpolants.py --cells 128; JP_ADVECT_HI_GLOBAL=1 timeout 300 python tools/time_interpolants.py --cells 128 ) 2>&1 | grep -E "RK2|checksum" | tee gpurun_out/r02s_interpolants.log
