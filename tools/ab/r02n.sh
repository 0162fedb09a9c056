set -x
mkdir -p gpurun_out
( for v in g_base g_pair; do
    echo "=== $v"; JUSTPIC_LIB=$PWD/tools/ab/libs/$v.so timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify 1 --interp 1 2>&1 | grep -E "^advect|checksum"
  done ) > gpurun_out/r02n_ab.log 2>&1
grep -E "===|^advect|checksum" gpurun_out/r02n_ab.log
