set -x
mkdir -p gpurun_out
timeout 600 python tools/all_kernels.py --cells 64 2>&1 | tail -3
timeout 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/all_128 python tools/all_kernels.py --cells 128 > gpurun_out/all_kernels_ncu.log 2>&1; tail -3 gpurun_out/all_kernels_ncu.log
ncu -i /tmp/all_128.ncu-rep --page raw --csv > gpurun_out/r01f_all_kernels_128_raw.csv
python profiles/kernel_table.py gpurun_out/r01f_all_kernels_128_raw.csv > gpurun_out/r01f_all_kernels_128_table.md; head -5 gpurun_out/r01f_all_kernels_128_table.md; wc -l gpurun_out/r01f_all_kernels_128_table.md
ls -la /tmp/all_128.ncu-rep
