set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m pytest tests/test_gpu_halo.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r01f_bench_2gpu.json 2> gpurun_out/r01f_bench_2gpu.err; tail -c 1500 gpurun_out/r01f_bench_2gpu.json; tail -5 gpurun_out/r01f_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 | tail -c 400
