# round 2, 2-GPU call (gpurun --gpus 2): every multi-GPU parity test incl. decomposed-vs-global-oracle over NCCL (0 skips expected),
# per-stage halo timing, weak-scaling bench with / without the overlap (state_checksum must agree), strong-scaling line
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
( time timeout 900 python -m pytest tests/test_gpu_halo.py tests/test_gpu_decomposed_vs_global.py -m gpu -q -rs 2>&1 | tail -40 ) > gpurun_out/r02p_pytest_2gpu.log 2>&1; cat gpurun_out/r02p_pytest_2gpu.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/time_halo.py 2>&1 | grep "^dim" | tee gpurun_out/r02p_time_halo.log
for o in 1 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$o bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --overlap $o > gpurun_out/r02p_bench_2gpu_overlap$o.json 2> gpurun_out/r02p_bench_2gpu_overlap$o.err
  tail -c 600 gpurun_out/r02p_bench_2gpu_overlap$o.json; tail -3 gpurun_out/r02p_bench_2gpu_overlap$o.err | cut -c1-300
done
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --scaling strong --global-cells 256 > gpurun_out/r02p_bench_2gpu_strong256.json 2> gpurun_out/r02p_bench_2gpu_strong256.err
tail -c 400 gpurun_out/r02p_bench_2gpu_strong256.json; grep -E "Init COMPLETE|nranks" gpurun_out/r02p_bench_2gpu_strong256.err | head -6 | cut -c1-250
python - <<'PY'
import json
for f in ("overlap1","overlap0","strong256"):
    try:
        d=json.loads(open(f"gpurun_out/r02p_bench_2gpu_{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["value"]/1e9,3), "G/s", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["phase_ms"].items()}, d["config"]["state_checksum"], d["config"]["topology"], d["config"]["migrant_fraction"])
    except Exception as e: print(f, "ERR", e)
PY
