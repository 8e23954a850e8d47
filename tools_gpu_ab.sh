#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 1 2 3; do
  B200_PDL=$pdl timeout 300 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-prefill > gpurun_out/ab_$pdl.json 2> gpurun_out/ab_$pdl.err
  python -c "
import json; j=json.load(open('gpurun_out/ab_$pdl.json')); print('PDL=$pdl', j['value'], j['ms_per_step'], j['roofline']['frac'])"
done
