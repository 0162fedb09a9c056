set -x
mkdir -p gpurun_out
( for v in pf4 pf2 pf3 pf4b3; do
    echo "=== $v"; JUSTPIC_LIB=$PWD/tools/ab/libs/$v.so timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify 1 --interp 1 2>&1 | grep -E "move stages|^move|^advect|checksum"
  done ) > gpurun_out/r02h_ab_gather.log 2>&1
grep -E "===|move stages|^move|^advect" gpurun_out/r02h_ab_gather.log
