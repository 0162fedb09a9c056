# 2-GPU checks (gpurun --gpus 2 -- 'bash tools/ab/run_2gpu.sh'): NCCL halo tests incl. the overlapped-path parity test,
# per-stage halo timing, bench with and without the overlap (state_checksum of the two lines must be identical)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_halo.py -m gpu -x -q 2>&1 | tail -5
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/time_halo.py 2>&1 | grep "^dim"
for o in 0 1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$o bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --overlap $o > gpurun_out/bench_2gpu_overlap$o.json 2> gpurun_out/bench_2gpu_overlap$o.err
  tail -c 400 gpurun_out/bench_2gpu_overlap$o.json
done
