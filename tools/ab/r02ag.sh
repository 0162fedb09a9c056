# prefetch of the neighbourhood's occupancy rows in the move plan (JP_PLAN_PREFETCH)
set -x
mkdir -p gpurun_out
for v in base ppf base ppf; do
  JUSTPIC_LIB=tools/ab/libs/$v.so python tools/time_phases.py --cells 256 --steps 6 --classify 1 --interp 1 2>&1 | grep -i "move stages\|^move" | tail -3 | sed "s/^/$v /"
  JUSTPIC_LIB=tools/ab/libs/$v.so python tools/time_phases.py --cells 128 --steps 6 --classify 1 --interp 1 2>&1 | grep -i "move stages" | tail -1 | sed "s/^/$v 128 /"
done | tee gpurun_out/r02ag_ab_plan_prefetch.log
