# memory-safety pass over the GPU parity tests (small grids): compute-sanitizer memcheck, then racecheck on the
# shared-memory kernels (tiled advection with / without the hand-off, fused phase ratios, move plan)
mkdir -p gpurun_out
( time timeout 700 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests -m gpu -x -q -k "trajectory or wide or handoff or force_injection or interpolations or inject or conversion or halo" 2>&1 | tail -25 ) > gpurun_out/sanitizer_memcheck.log 2>&1
tail -30 gpurun_out/sanitizer_memcheck.log
( time timeout 500 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_handoff.py tests/test_phase_ratios_nodes.py -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/sanitizer_racecheck.log 2>&1
tail -30 gpurun_out/sanitizer_racecheck.log
