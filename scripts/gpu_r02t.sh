#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02t_launches_variants.csv \
    python bench.py --workload variants --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02t_ncu_variants.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r02t_launches_variants.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki].split('(')[0][-50:]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
for k,(n,v) in agg.items(): print(f"{v/1e6:9.3f} ms {100*v/tot:5.1f}% x{n:4d} {k}")
print('total', tot/1e6)
PY
