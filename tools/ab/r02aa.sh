# CUDA-graph capture of a whole step: parity of the replays, and the launch-bound configurations eager vs replayed
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py -q -x 2>&1 | tail -15 | tee gpurun_out/r02aa_pytest_graph.log
for c in cfg1 cfg2 cfg3; do
  timeout 300 python bench.py --config $c --steps 20 --graph 1 > gpurun_out/r02aa_bench_${c}_graph.json 2> gpurun_out/r02aa_${c}.err || tail -5 gpurun_out/r02aa_${c}.err
  python - <<PY
import json
l=[x for x in open("gpurun_out/r02aa_bench_${c}_graph.json") if x.startswith('{"metric"')]
if l:
    d=json.loads(l[-1]); print("${c}", "graph", d["ms_per_step"], "eager", d["config"]["eager_ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], d["phase_ms"])
PY
done
