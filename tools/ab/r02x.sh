# 'bucketed' state: first-interpolation fast path only when the particles are known to sit in their cells
set -x
mkdir -p gpurun_out
python tools/time_interpolants.py --cells 128 > gpurun_out/r02x_interpolants_128.log 2>&1; cat gpurun_out/r02x_interpolants_128.log
python bench.py > gpurun_out/r02x_bench_256.json 2> gpurun_out/r02x_bench_256.err; cat gpurun_out/r02x_bench_256.json
timeout 900 python -m pytest tests -q -m gpu -x -k "advect or trajectory or handoff or interp" 2>&1 | tail -3 | tee gpurun_out/r02x_pytest_subset.log
