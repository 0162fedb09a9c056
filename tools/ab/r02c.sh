# round 2, call c: A/B of the fused scatter variants (256^3, 8 steps each), failed halo test
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_halo.py tests/test_gpu_decomposed_vs_global.py tests/test_gpu_interp_handoff.py -m gpu -x -q 2>&1 | tail -5
( for v in default u2 u2reg u4reg u4c u4c3 u2c3; do
    lib=tools/ab/libs/$v.so; [ $v = default ] && lib=justpic/jl_b200/libjustpic_sm100a.so
    echo "=== $v"; JUSTPIC_LIB=$PWD/$lib timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify 1 --interp 1 2>&1 | grep -E "move stages|^move|^advect|checksum"
  done
  echo "=== generic"; JP_SCI_GENERIC=1 timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify 1 --interp 1 2>&1 | grep -E "move stages|^move|checksum"
) > gpurun_out/r02c_ab_scatter.log 2>&1
cat gpurun_out/r02c_ab_scatter.log
