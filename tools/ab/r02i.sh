set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_advect_tile|k_move_scatter_interp|k_move_gather|k_p2g_node|k_move_finalize' --launch-skip 15 --launch-count 5 -f -o gpurun_out/r02i_full_256 python tools/time_phases.py --cells 256 --steps 5 --classify 1 --interp 1 > gpurun_out/r02i_ncu.log 2>&1; tail -3 gpurun_out/r02i_ncu.log
ncu -i gpurun_out/r02i_full_256.ncu-rep --page raw --csv > gpurun_out/r02i_ncu_full_256_raw.csv
ncu -i gpurun_out/r02i_full_256.ncu-rep --page source --csv --kernel-name regex:k_move_scatter_interp > gpurun_out/r02i_scatter_source.csv
ls -la gpurun_out | grep r02i
