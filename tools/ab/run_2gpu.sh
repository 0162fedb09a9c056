mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo.py -m gpu -x -q 2>&1 | grep -v Warning | tail -40 | cut -c1-220
run() { # $1 tag, $2 overlap, extra env already exported
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --overlap $2 > gpurun_out/ovl_$1.json 2> gpurun_out/ovl_$1.err; tail -2 gpurun_out/ovl_$1.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/ovl_$1.json').read().strip().splitlines()[-1])
print("$1 overlap", $2, round(d['value']/1e9,3), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phase_ms'].items()})
PY
}
run seq 0
run ovl256 1
NCCL_NTHREADS=512 run ovl512 1
NCCL_NTHREADS=128 run ovl128 1
NCCL_NTHREADS=256 NCCL_MAX_NCHANNELS=8 run ovl256c8 1
