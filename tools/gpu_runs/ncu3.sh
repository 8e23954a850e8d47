#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 1100 --csv --log-file gpurun_out/r01d_launches_prefill.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench5.log 2>&1
python - <<'PY'
import csv, collections, re
with open('gpurun_out/r01d_launches_prefill.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.defaultdict(lambda:[0,0.0]); tot=0
for row in r:
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=re.sub(r'\(.*','',row['Kernel Name'])[:70]
    v=float(row['Metric Value'].replace(',','')); unit=row['Metric Unit']
    if unit=='ns': v/=1000
    elif unit=='ms': v*=1000
    agg[name][0]+=1; agg[name][1]+=v; tot+=v
print("total us", round(tot,1), "launches", sum(a[0] for a in agg.values()))
for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d} avg={t/n:8.2f}us  {k}")
PY
