# final build, 2-GPU call (gpurun --gpus 2): every multi-GPU parity test (0 skips expected), weak-scaling bench with / without the overlap
set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_decomposed_vs_global.py -m gpu -q -rs 2>&1 | tail -12 ) > gpurun_out/r02ac_pytest_2gpu.log 2>&1; cat gpurun_out/r02ac_pytest_2gpu.log
for o in 1 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$o bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --overlap $o > gpurun_out/r02ac_bench_2gpu_overlap$o.json 2> gpurun_out/r02ac_bench_2gpu_overlap$o.err
  tail -3 gpurun_out/r02ac_bench_2gpu_overlap$o.err | cut -c1-300
done
python - <<'PY'
import json
for f in ("overlap1","overlap0"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02ac_bench_2gpu_{f}.json") if l.startswith('{"metric"')][-1])
        print(f, round(d["value"]/1e9,3), "G/s", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["phase_ms"].items()}, d["config"]["state_checksum"], d["config"]["topology"], d["gpu_launches"])
    except Exception as e: print(f, "ERR", e)
PY
