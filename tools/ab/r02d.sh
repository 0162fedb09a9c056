set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_interp_handoff.py -m gpu -x -q 2>&1 | tail -60 > gpurun_out/r02d_pytest.log; cat gpurun_out/r02d_pytest.log
