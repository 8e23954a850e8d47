#!/bin/bash
mkdir -p gpurun_out
for cfg in "3 0" "3 1" "1 1" "0 0"; do
  set -- $cfg
  B200_PDL=$1 B200_PDL_EARLY=$2 timeout 300 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-prefill > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; j=json.load(open('gpurun_out/ab.json')); print('PDL=$1 early=$2', j['value'], j['ms_per_step'], j['roofline']['frac'])"
done
