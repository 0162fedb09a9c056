# the three slot policies of move_particles! side by side (reference = credited; compact / dense = opt-in), at a settled state (20 warm-up steps)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "policy" 2>&1 | tail -3 | tee gpurun_out/r02y_pytest_policy.log
for pol in reference compact dense; do
  python bench.py --warmup 20 --steps 10 --move-policy $pol --no-cpu-baseline > gpurun_out/r02y_bench_256_${pol}_w20.json 2> gpurun_out/r02y_${pol}.err
  python - <<PY
import json
l=[x for x in open("gpurun_out/r02y_bench_256_${pol}_w20.json") if x.startswith('{"metric"')]
if not l: print(open("gpurun_out/r02y_${pol}.err").read()[-2000:])
else:
    d=json.loads(l[-1]); c=d["config"]
    print("${pol}", d["ms_per_step"], d["value"], d["phase_ms"], d["move_stage_ms"], c["slot_fill"], c["dropped_per_step"], c["move_path"], d["roofline"]["frac"], d["roofline"]["step_frac"])
PY
done
