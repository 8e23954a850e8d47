#!/bin/bash
# r01h: launch list of two decode steps in the final schedule (7 kernels/layer, PDL) + full capture of the w13 mat-vec
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2300 -c 460 --csv --log-file gpurun_out/r01h_launches_decode.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prefill > gpurun_out/ncu_bench10.log 2>&1
python - <<'PY'
import csv, collections, re
with open('gpurun_out/r01h_launches_decode.csv') as f:
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
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mmv_fused -s 40 -c 4 -f -o gpurun_out/r01h_mmv_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prefill > gpurun_out/ncu_bench11.log 2>&1
ls -la gpurun_out/r01h_mmv_fused.ncu-rep
