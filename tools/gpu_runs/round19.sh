#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/gpt2_debug.py gpt2 2>&1 | tail -6
for i in 1 2; do
( timeout 1200 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider ) > gpurun_out/test_gpu_full_$i.log 2>&1
grep -E "AssertionError|passed|failed|FAILED" gpurun_out/test_gpu_full_$i.log | head -8
done
timeout 300 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err
python -c "
import json; j=json.load(open('gpurun_out/ab.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['prefill']['ms'])"
