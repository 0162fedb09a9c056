# usage: bash tools/ab/r02q_ngpu.sh N   (inside gpurun --gpus N): weak scaling (default = what the driver runs), without overlap, strong 512^3
N=$1
mkdir -p gpurun_out
run() { # name, extra args
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline $2 > gpurun_out/r02q_bench_${N}gpu_$1.json 2> gpurun_out/r02q_bench_${N}gpu_$1.err
  tail -2 gpurun_out/r02q_bench_${N}gpu_$1.err | cut -c1-200
  python - <<PY
import json
try:
    d=[json.loads(l) for l in open('gpurun_out/r02q_bench_${N}gpu_$1.json') if l.startswith('{"metric"')][-1]
    print("N=$N $1", round(d['value']/1e9,3), "G/s", round(d['ms_per_step'],2), "ms", {k:round(v,2) for k,v in d['phase_ms'].items()}, "e2e", round(d['e2e']['value']/1e9,3) if 'e2e' in d else None, d['config']['topology'], d['config']['block_cells'], d['config']['migrant_fraction'], d['config']['state_checksum'])
except Exception as e: print("N=$N $1 ERR", e)
PY
}
run weak ""

run strong512 "--scaling strong --global-cells 512 --no-e2e"

true
