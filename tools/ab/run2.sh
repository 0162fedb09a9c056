mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for v in old vec vec4 old vec; do
  echo "=== $v"; JUSTPIC_LIB=$PWD/tools/ab/libs/$v.so timeout 200 python tools/time_phases.py --cells 256 --steps 6 --classify 1 2>&1 | tail -6
done > gpurun_out/ab_scatter_vec.log 2>&1
grep "===\|move\|checksum" gpurun_out/ab_scatter_vec.log
