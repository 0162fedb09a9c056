# FINAL build, 4-GPU call (gpurun --gpus 4): weak-scaling bench line (256^3 per GPU, topology 1x2x2, overlap on)
set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02aj_bench_4gpu.json 2> gpurun_out/r02aj_bench_4gpu.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02aj_bench_4gpu.json") if l.startswith('{"metric"')][-1])
print(round(d["value"]/1e9,3), "G/s", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["phase_ms"].items()}, d["config"]["state_checksum"], d["config"]["topology"], d["gpu_launches"], "e2e", round(d["e2e"]["value"]/1e9,2))
PY
tail -3 gpurun_out/r02aj_bench_4gpu.err | cut -c1-300
