# round 2, call f: whole GPU suite incl. the full-size parity at all four BASELINE sizes, bench (headline + cfg1-3 + reference arm), launch list
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=6 2>&1 | tail -25 ) > gpurun_out/r02f_pytest_gpu.log 2>&1; cat gpurun_out/r02f_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02f_bench_256.json 2> gpurun_out/r02f_bench.err; tail -c 1200 gpurun_out/r02f_bench_256.json; tail -3 gpurun_out/r02f_bench.err
for c in cfg1 cfg2 cfg3; do timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r02f_bench_$c.json 2> gpurun_out/r02f_bench_$c.err; tail -c 900 gpurun_out/r02f_bench_$c.json; tail -2 gpurun_out/r02f_bench_$c.err; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2>> gpurun_out/r02f_bench.err; cat gpurun_out/r02f_bench_reference_arm.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_launches_256.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02f_launch_bench.log 2>&1; tail -2 gpurun_out/r02f_launch_bench.log | cut -c1-200; wc -l gpurun_out/r02f_launches_256.csv
