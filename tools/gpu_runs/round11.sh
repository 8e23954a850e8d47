#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider ) > gpurun_out/test_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_full.log
grep -E "AssertionError|passed|failed|FAILED|real" gpurun_out/test_gpu_full.log | head -20
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench.json'))
print({k:j[k] for k in ('value','ms_per_step','launches_per_step')}, j['roofline']['frac'], j['step_roofline']['frac'], j['e2e']['value'])
print(j.get('prefill')); print(j.get('fast_mode'))
PY
tail -4 gpurun_out/bench.err
