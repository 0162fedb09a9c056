set -x
mkdir -p gpurun_out
for v in v0 v1 v2 v3 v4 v5 v0; do
  echo "=== $v"; JUSTPIC_LIB=$PWD/tools/ab/libs/$v.so timeout 200 python tools/time_phases.py --cells 256 --steps 6 --classify 1 2>&1 | tail -7
done > gpurun_out/ab_shapes.log 2>&1
cat gpurun_out/ab_shapes.log | grep "===\|advect\|checksum"
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 420 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/all_128 python tools/all_kernels.py --cells 128 > gpurun_out/all_kernels_ncu.log 2>&1; tail -3 gpurun_out/all_kernels_ncu.log
ncu -i /tmp/all_128.ncu-rep --page raw --csv > gpurun_out/r01f_all_kernels_128_raw.csv
python profiles/kernel_table.py gpurun_out/r01f_all_kernels_128_raw.csv > gpurun_out/r01f_all_kernels_128_table.md; wc -l gpurun_out/r01f_all_kernels_128_table.md
