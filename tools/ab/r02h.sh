set -x
mkdir -p gpurun_out
( for v in base pipe pipesm pipesmc; do
    echo "=== $v"; JUSTPIC_LIB=$PWD/tools/ab/libs/$v.so timeout 300 python tools/time_phases.py --cells 256 --steps 8 --classify 1 --interp 1 2>&1 | grep -E "move stages|^move|^advect|checksum"
  done ) > gpurun_out/r02j_ab_pipe.log 2>&1
grep -E "===|move stages|^move|^advect" gpurun_out/r02j_ab_pipe.log
