#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider ) > gpurun_out/test_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_full.log
grep -E "AssertionError|passed|failed|FAILED|real" gpurun_out/test_gpu_full.log | head -20
for cfg in 4 6 8 12; do
  B200_RING_DEPTH=$cfg timeout 300 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-prefill > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json; j=json.load(open('gpurun_out/ab.json')); print('depth=$cfg', j['value'], j['ms_per_step'], j['launches_per_step'], j['roofline']['frac'])"
done
