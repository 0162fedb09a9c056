# round-1 final evidence run (one gpurun call): GPU parity tests, bench (both arms), launch list,
# per-kernel metric table over EVERY kernel, one --set full capture of the headline step's kernels.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
( time timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/r01f_pytest_gpu.log 2>&1; cat gpurun_out/r01f_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r01f_bench_256.json 2> gpurun_out/r01f_bench.err; tail -c 600 gpurun_out/r01f_bench_256.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01f_bench_reference_arm.json 2>> gpurun_out/r01f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01f_launches_256.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r01f_launch_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 420 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/all_128 python tools/all_kernels.py --cells 128 > gpurun_out/all_kernels_ncu.log 2>&1; tail -3 gpurun_out/all_kernels_ncu.log
ncu -i /tmp/all_128.ncu-rep --page raw --csv > gpurun_out/r01f_all_kernels_128_raw.csv
python profiles/kernel_table.py gpurun_out/r01f_all_kernels_128_raw.csv > gpurun_out/r01f_all_kernels_128_table.md; head -5 gpurun_out/r01f_all_kernels_128_table.md; wc -l gpurun_out/r01f_all_kernels_128_table.md
# --set full of the big kernels of the hand-off step at 256^3 (4th step)
timeout 480 ncu --set full --clock-control none --import-source on -k regex:'k_advect_tile|k_move_gather|k_move_scatter|k_p2g_cell|k_phase' --launch-skip 15 --launch-count 5 -f -o gpurun_out/r01f_full_256 python tools/time_phases.py --cells 256 --steps 5 --classify 1 > gpurun_out/r01f_full_256.log 2>&1; tail -3 gpurun_out/r01f_full_256.log
ncu -i gpurun_out/r01f_full_256.ncu-rep --page raw --csv > gpurun_out/r01f_ncu_full_256_raw.csv
ls -la gpurun_out
