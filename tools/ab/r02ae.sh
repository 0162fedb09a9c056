# compile-time 'bucketed' variants of k_advect_tile: parity subset, second-advection timing, headline
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -k "advect or trajectory or handoff or graph or halo or interp" 2>&1 | tail -3 | tee gpurun_out/r02ae_pytest_subset.log
python tools/time_interpolants.py --cells 128 2>&1 | grep -v Warn | tee gpurun_out/r02ae_interpolants_128.log
python bench.py --no-cpu-baseline > gpurun_out/r02ae_bench_256.json 2> gpurun_out/r02ae_256.err
python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/r02ae_bench_256.json") if x.startswith('{"metric"')][-1])
print("256^3", d["ms_per_step"], d["value"], d["phase_ms"], d["move_stage_ms"], d["gpu_launches"], d["e2e"]["value"])
PY
