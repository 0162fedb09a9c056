# FINAL build: ncu --set full of the three big kernels of the 256^3 step (one launch each, 3rd step)
set -x
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_advect_tile|k_move_scatter_interp|k_move_gather' --launch-skip 6 --launch-count 3 -f -o /tmp/r02al_full_256 python tools/time_phases.py --cells 256 --steps 4 --classify 1 --interp 1 > gpurun_out/r02al_ncu.log 2>&1; tail -2 gpurun_out/r02al_ncu.log
ncu -i /tmp/r02al_full_256.ncu-rep --page raw --csv > gpurun_out/r02al_ncu_full_256_raw.csv
python profiles/ncu_summary.py gpurun_out/r02al_ncu_full_256_raw.csv > gpurun_out/r02al_ncu_full_256_summary.txt 2>&1; head -12 gpurun_out/r02al_ncu_full_256_summary.txt
