set -x
mkdir -p gpurun_out
( for v in d_base d_pad1 d_pad1ex40 d_pad3; do
    echo "=== $v"; JUSTPIC_LIB=$PWD/tools/ab/libs/$v.so timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify 1 --interp 1 2>&1 | grep -E "^advect|checksum"
  done ) > gpurun_out/r02l_ab_advect.log 2>&1
grep -E "===|^advect|checksum" gpurun_out/r02l_ab_advect.log
