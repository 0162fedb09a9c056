# usage: bash tools/ab/run_ngpu.sh N   (inside gpurun --gpus N)
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_bench_${N}gpu.json 2> gpurun_out/r01f_bench_${N}gpu.err
tail -c 300 gpurun_out/r01f_bench_${N}gpu.json; tail -3 gpurun_out/r01f_bench_${N}gpu.err
