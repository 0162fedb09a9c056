set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "inject or trajectory or rotating or cfg1 or cfg2 or cfg3 or wide" 2>&1 | tail -4
for c in cfg3 cfg1 cfg2; do timeout 600 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r02t_bench_$c.json 2> gpurun_out/r02t_bench_$c.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02t_bench_$c.json').read().strip().splitlines()[-1])
print("$c", round(d["value"]/1e9,3), "G/s", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["phase_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e9,3))
PY
done
