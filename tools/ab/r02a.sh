# round 2, first call: full-size GPU-vs-oracle parity (cfg1-cfg4), the whole GPU suite, baseline bench
set -x
mkdir -p gpurun_out
nproc; free -g | head -2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
( time timeout 900 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -x -q --durations=5 2>&1 | tail -15 ) > gpurun_out/r02a_fullsize.log 2>&1; cat gpurun_out/r02a_fullsize.log
( time timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize_parity.py 2>&1 | tail -5 ) > gpurun_out/r02a_pytest_gpu.log 2>&1; cat gpurun_out/r02a_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02a_bench_256.json 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench_256.json
