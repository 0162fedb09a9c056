set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_256_r01e.json 2> gpurun_out/bench_256_r01e.err
tail -c 300 gpurun_out/bench_256_r01e.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_advect_tile|k_move_classify3|k_move_gather|k_move_scatter|k_p2g_cell|k_phase" --launch-skip 18 -c 6 -o gpurun_out/prof_256_r01e python tools/time_phases.py --cells 256 --steps 5 > gpurun_out/ncu_256_r01e.log 2>&1
ncu -i gpurun_out/prof_256_r01e.ncu-rep --page raw --csv > gpurun_out/prof_256_r01e_raw.csv
python profiles/ncu_summary.py gpurun_out/prof_256_r01e_raw.csv > gpurun_out/prof_256_r01e_summary.txt
grep -E "Kernel Name|gpu__time_duration|dram__bytes" gpurun_out/prof_256_r01e_summary.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01e.json 2>gpurun_out/bench_ref_r01e.err; cat gpurun_out/bench_ref_r01e.json | cut -c1-400
