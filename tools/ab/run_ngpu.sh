# usage: bash tools/ab/run_ngpu.sh N   (inside gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_halo.py -m gpu -x -q -k "two_gpus" 2>&1 | tail -3
for o in 1 0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$o bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --overlap $o > gpurun_out/r01g_bench_${N}gpu_ovl$o.json 2> gpurun_out/r01g_bench_${N}gpu_ovl$o.err
tail -2 gpurun_out/r01g_bench_${N}gpu_ovl$o.err | cut -c1-200
python - <<PY
import json
d=json.loads(open('gpurun_out/r01g_bench_${N}gpu_ovl$o.json').read().strip().splitlines()[-1])
print("N=$N overlap", $o, round(d['value']/1e9,3), round(d['ms_per_step'],2), {k:round(v,2) for k,v in d['phase_ms'].items()}, "e2e", round(d['e2e']['value']/1e9,3))
PY
done
