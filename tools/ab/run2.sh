# A/B of compile-time variants (tools/ab/libs/*.so) on the headline step at 256^3, hand-off on
mkdir -p gpurun_out
for v in base ex48 ex40 occhi occlo u8 u2 base; do
  echo "=== $v"; JUSTPIC_LIB=$PWD/tools/ab/libs/$v.so timeout 200 python tools/time_phases.py --cells 256 --steps 6 --classify 1 2>&1 | tail -6
done > gpurun_out/ab_variants.log 2>&1
grep "===\|advect\|move\|p2g\|checksum" gpurun_out/ab_variants.log
