set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu_wide.log 2>&1; cat gpurun_out/pytest_gpu_wide.log
