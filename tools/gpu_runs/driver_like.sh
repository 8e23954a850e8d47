#!/bin/bash
# what the driver runs at round end: serial GPU tests, smoke(), the default bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/test_gpu_serial.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|real" gpurun_out/test_gpu_serial.log | head -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
( time timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench_full.json'))
print({k:j[k] for k in ('value','ms_per_step','launches_per_step','steps','warmup')}, 'roofline', round(j['roofline']['frac'],4), 'step', round(j['step_roofline']['frac'],4), 'e2e', j['e2e']['value'])
print('prefill', j['prefill']['ms'], 'cpu', j['cpu_baseline']['value'], j['cpu_baseline']['cores'], 'clocks', j['clocks'])
PY
