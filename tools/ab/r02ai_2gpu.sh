# FINAL build, 2-GPU call (gpurun --gpus 2): every multi-GPU parity test, weak-scaling bench with the overlap (the split advection keeps the bucketed state)
set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_halo.py tests/test_gpu_decomposed_vs_global.py -m gpu -q -rs 2>&1 | tail -6 ) > gpurun_out/r02ai_pytest_2gpu.log 2>&1; cat gpurun_out/r02ai_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02ai_bench_2gpu.json 2> gpurun_out/r02ai_bench_2gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02ai_bench_2gpu.json") if l.startswith('{"metric"')][-1])
print(round(d["value"]/1e9,3), "G/s", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["phase_ms"].items()}, d["config"]["state_checksum"], d["config"]["topology"], d["gpu_launches"], "e2e", round(d["e2e"]["value"]/1e9,2))
PY
