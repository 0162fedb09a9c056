# round 2, final build, 1 GPU: whole suite incl. full-size parity, smoke, bench (+ hand-offs off, + reference arm), launch list
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=4 2>&1 | tail -12 ) > gpurun_out/r02z_pytest_gpu.log 2>&1; cat gpurun_out/r02z_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02z_bench_256.json 2> gpurun_out/r02z_bench.err; tail -3 gpurun_out/r02z_bench.err
timeout 300 python bench.py --handoff 0 --interp-handoff 0 --no-cpu-baseline --no-e2e > gpurun_out/r02z_bench_256_nohandoffs.json 2>> gpurun_out/r02z_bench.err
python - <<'PY'
import json
for f in ("r02z_bench_256","r02z_bench_256_nohandoffs"):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, round(d["value"]/1e9,3), "G/s", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["phase_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e9,3) if "e2e" in d else None, {k:round(v,3) for k,v in d["move_stage_ms"].items()}, "frac", round(d["roofline"]["frac"],3), round(d["roofline"]["step_frac"],3))
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02z_bench_reference_arm.json 2>> gpurun_out/r02z_bench.err; cat gpurun_out/r02z_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches_256.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02z_launch_bench.log 2>&1; wc -l gpurun_out/r02z_launches_256.csv
