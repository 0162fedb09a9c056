# evict-first hints on single-use data in gather / fused scatter (JP_STREAM_HINTS)
set -x
mkdir -p gpurun_out
for v in base cs base cs; do
  JUSTPIC_LIB=tools/ab/libs/$v.so python tools/time_phases.py --cells 256 --steps 6 --classify 1 --interp 1 2>&1 | grep -i "move stages\|^move\|step " | tail -3 | sed "s/^/$v /"
done | tee gpurun_out/r02af_ab_stream_hints.log
