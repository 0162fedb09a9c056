# per-kernel metric table over EVERY kernel of the library at 128^3 (round-2 build)
set -x
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/all_128 python tools/all_kernels.py --cells 128 > gpurun_out/r02w_all_kernels_ncu.log 2>&1; tail -3 gpurun_out/r02w_all_kernels_ncu.log
ncu -i /tmp/all_128.ncu-rep --page raw --csv > gpurun_out/r02w_all_kernels_128_raw.csv
python profiles/kernel_table.py gpurun_out/r02w_all_kernels_128_raw.csv > gpurun_out/r02w_all_kernels_128_table.md; head -5 gpurun_out/r02w_all_kernels_128_table.md; wc -l gpurun_out/r02w_all_kernels_128_table.md
