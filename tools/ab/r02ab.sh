# cooperative plan / direct-sweep kernels (3^N launches -> 1 each): parity subset + the four configurations
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -k "move or trajectory or handoff or graph or policy or ties or wide or halo or inject" 2>&1 | tail -4 | tee gpurun_out/r02ab_pytest_subset.log
for c in cfg1 cfg2 cfg3; do
  timeout 300 python bench.py --config $c --steps 20 --graph 1 > gpurun_out/r02ab_bench_${c}_graph.json 2> gpurun_out/r02ab_${c}.err || tail -5 gpurun_out/r02ab_${c}.err
  python - <<PY
import json
l=[x for x in open("gpurun_out/r02ab_bench_${c}_graph.json") if x.startswith('{"metric"')]
if l:
    d=json.loads(l[-1]); print("${c}", "graph", d["ms_per_step"], "eager", d["config"]["eager_ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], d["phase_ms"])
PY
done
python bench.py --no-cpu-baseline > gpurun_out/r02ab_bench_256.json 2> gpurun_out/r02ab_256.err
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/r02ab_bench_256.json") if x.startswith('{"metric"')][-1])
print("256^3", d["ms_per_step"], d["value"], d["phase_ms"], d["move_stage_ms"], d["gpu_launches"], d["e2e"]["value"])
PY
