# round 2, call b: new hand-off / halo / decomposed tests, A/B of the move -> interpolation hand-off, bench, ncu source pages
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize_parity.py 2>&1 | tail -15 ) > gpurun_out/r02b_pytest_gpu.log 2>&1; cat gpurun_out/r02b_pytest_gpu.log
( time timeout 300 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -x -q -k "cfg1 or cfg2" 2>&1 | tail -5 ) > gpurun_out/r02b_fullsize12.log 2>&1; cat gpurun_out/r02b_fullsize12.log
for i in 0 1; do timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify 1 --interp $i 2>&1 | tail -8; done > gpurun_out/r02b_ab_interp.log 2>&1; cat gpurun_out/r02b_ab_interp.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02b_bench_256.json 2> gpurun_out/r02b_bench.err; tail -c 2500 gpurun_out/r02b_bench_256.json; tail -5 gpurun_out/r02b_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_advect_tile|k_move_scatter_interp|k_move_gather' --launch-skip 9 --launch-count 3 -f -o gpurun_out/r02b_full_256 python tools/time_phases.py --cells 256 --steps 4 --classify 1 --interp 1 > gpurun_out/r02b_ncu.log 2>&1; tail -3 gpurun_out/r02b_ncu.log
ncu -i gpurun_out/r02b_full_256.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_full_256_raw.csv
ls -la gpurun_out | tail -12
