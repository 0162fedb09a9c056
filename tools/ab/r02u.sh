set -x
mkdir -p gpurun_out
for v in i2 i3 i4; do JUSTPIC_LIB=$PWD/tools/ab/libs/$v.so timeout 600 python bench.py --config cfg3 --steps 10 --warmup 4 > gpurun_out/r02u_$v.json 2> gpurun_out/r02u_$v.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02u_$v.json').read().strip().splitlines()[-1])
print("$v", round(d["ms_per_step"],3), "ms", {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done
