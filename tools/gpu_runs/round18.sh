#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider ) > gpurun_out/test_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_full.log
grep -E "AssertionError|passed|failed|FAILED|real" gpurun_out/test_gpu_full.log | head -20
python tools/decode_timeline.py 32 2>&1 | tail -12
timeout 300 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-prefill > gpurun_out/ab.json 2> gpurun_out/ab.err
python -c "
import json; j=json.load(open('gpurun_out/ab.json')); print(j['value'], j['ms_per_step'], j['launches_per_step'], j['roofline']['frac'], j['e2e']['value'])"
