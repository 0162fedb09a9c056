# what does the runtime 'bucketed' flag cost in k_advect_tile? (adv_rt = flag read at run time, adv_ct = assumed at compile time)
set -x
mkdir -p gpurun_out
for v in adv_rt adv_ct adv_rt adv_ct; do
  JUSTPIC_LIB=tools/ab/libs/$v.so python tools/time_phases.py --cells 256 --steps 6 --classify 1 --interp 1 2>&1 | grep -i "advect\|step" | tail -3 | sed "s/^/$v /"
done | tee gpurun_out/r02ad_ab_bucketed_flag.log
