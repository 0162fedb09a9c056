# round 2, FINAL build, 1 GPU: whole suite incl. full-size parity, smoke, bench (+ hand-offs off, + reference arm), launch list, per-kernel table
set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=4 2>&1 | tail -12 ) > gpurun_out/r02ah_pytest_gpu.log 2>&1; cat gpurun_out/r02ah_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r02ah_bench_256.json 2> gpurun_out/r02ah_bench.err; tail -3 gpurun_out/r02ah_bench.err
timeout 300 python bench.py --handoff 0 --interp-handoff 0 --no-cpu-baseline --no-e2e > gpurun_out/r02ah_bench_256_nohandoffs.json 2>> gpurun_out/r02ah_bench.err
python - <<'PY'
import json
for f in ("r02ah_bench_256","r02ah_bench_256_nohandoffs"):
    d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{"metric"')][-1])
    print(f, round(d["value"]/1e9,3), "G/s", round(d["ms_per_step"],2), "ms", {k:round(v,2) for k,v in d["phase_ms"].items()}, "e2e", round(d["e2e"]["value"]/1e9,3) if "e2e" in d else None, {k:round(v,3) for k,v in d["move_stage_ms"].items()}, "frac", round(d["roofline"]["frac"],3), round(d["roofline"]["step_frac"],3), d["gpu_launches"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02ah_bench_reference_arm.json 2>> gpurun_out/r02ah_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02ah_launches_256.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02ah_launch_bench.log 2>&1; wc -l gpurun_out/r02ah_launches_256.csv
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__throughput.avg.pct_of_peak_sustained_elapsed
timeout 600 ncu --metrics $M --clock-control none --profile-from-start off -f -o /tmp/all_128 python tools/all_kernels.py --cells 128 > gpurun_out/r02ah_all_kernels_ncu.log 2>&1; tail -3 gpurun_out/r02ah_all_kernels_ncu.log
ncu -i /tmp/all_128.ncu-rep --page raw --csv > gpurun_out/r02ah_all_kernels_128_raw.csv
python profiles/kernel_table.py gpurun_out/r02ah_all_kernels_128_raw.csv > gpurun_out/r02ah_all_kernels_128_table.md; wc -l gpurun_out/r02ah_all_kernels_128_table.md
