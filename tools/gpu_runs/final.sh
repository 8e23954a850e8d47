#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider ) > gpurun_out/test_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_full.log
grep -E "AssertionError|passed|failed|FAILED|real" gpurun_out/test_gpu_full.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_full.json'))
print({k:j[k] for k in ('value','ms_per_step','launches_per_step','steps','warmup')}, 'roofline', round(j['roofline']['frac'],4), 'step', round(j['step_roofline']['frac'],4), 'e2e', j['e2e']['value'])
print('prefill', j['prefill']['ms'], 'cpu', j.get('cpu_baseline'), 'clocks', j['clocks'])
PY
tail -3 gpurun_out/bench_full.err
