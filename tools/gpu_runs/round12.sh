#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider ) > gpurun_out/test_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_full.log
grep -E "AssertionError|passed|failed|FAILED|real" gpurun_out/test_gpu_full.log | head -20
for pdl in 0 1; do
  B200_PDL=$pdl timeout 300 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/ab_$pdl.json 2> gpurun_out/ab_$pdl.err
  python -c "
import json; j=json.load(open('gpurun_out/ab_$pdl.json')); print('PDL=$pdl', j['value'], j['ms_per_step'], j['roofline']['frac'], j['prefill']['ms'])"
done
