set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -k "not cfg4" 2>&1 | tail -8 ) > gpurun_out/r02g_pytest_gpu.log 2>&1; cat gpurun_out/r02g_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02g_bench_256.json 2> gpurun_out/r02g_bench.err; tail -3 gpurun_out/r02g_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02g_bench_256.json').read().strip().splitlines()[-1])
print("headline", round(d["value"]/1e9,3), "G/s", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["phase_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e9,3), {k:round(v,3) for k,v in d["move_stage_ms"].items()})
PY
