# memory-safety pass over the GPU parity tests that exercise the round-2 kernels (small grids): compute-sanitizer memcheck, then
# racecheck on the shared-memory kernels (tiled advection with the row-wise byte table, fused phase ratios), then initcheck-free synccheck
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests -m gpu -x -q -k "interp_handoff or decomposed_on_one or self_wrap or pack_unpack or handoff_trajectory or invalidate or trajectory_advect_move_inject or staging" 2>&1 | tail -25 ) > gpurun_out/r02r_sanitizer_memcheck.log 2>&1
tail -12 gpurun_out/r02r_sanitizer_memcheck.log
( time timeout 600 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_handoff.py tests/test_gpu_interp_handoff.py -m gpu -x -q -k "trajectory or equals_standalone" 2>&1 | tail -25 ) > gpurun_out/r02r_sanitizer_racecheck.log 2>&1
tail -12 gpurun_out/r02r_sanitizer_racecheck.log
