set -x
mkdir -p gpurun_out
JUSTPIC_LIB=$PWD/tools/ab/libs/i4.so timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_inject' -c 40 --csv --log-file gpurun_out/r02v_inject_launches.csv python bench.py --config cfg3 --steps 2 --warmup 2 > gpurun_out/r02v.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r02v_inject_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value'); idc=h.index('ID')
per=collections.OrderedDict()
for r in rows[hdr+1:]:
    if len(r)<=mv: continue
    per.setdefault((r[idc], r[kn].split('(')[0][:40]), {})[r[mn]]=float(r[mv].replace(',',''))
for (i,k),m in list(per.items())[-18:]:
    print(i, k, {a.split('__')[-1][:22]: round(b/1e6,3) if 'time' in a or 'inst' in a else round(b/1e9,3) for a,b in m.items()})
PY
