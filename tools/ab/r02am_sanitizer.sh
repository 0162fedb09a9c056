# compute-sanitizer memcheck over the tests of what the last third of round 2 added: dense policy (k_move_prevacate), cooperative plan / sweep
# kernels (auto and direct move), CUDA-graph capture, LinP / MQS tiles
mkdir -p gpurun_out
( time timeout 150 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests -m gpu -x -q -k "policy_dense or (trajectory_advect_move_inject and 3D-10) or (captured_step and 2D-24) or (advection_interpolants and 3D)" 2>&1 | tail -12 ) > gpurun_out/r02am_sanitizer_memcheck.log 2>&1
tail -12 gpurun_out/r02am_sanitizer_memcheck.log
