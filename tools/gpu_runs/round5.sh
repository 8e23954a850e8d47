#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 --timeout 600 -p no:cacheprovider -k "decode_kernel or 7b_geometry or native_vs_oracle or stream" > gpurun_out/test_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu.log
grep -E "AssertionError|passed|failed|FAILED|Error|error" gpurun_out/test_gpu.log | head -20
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --no-prefill > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/bench.json'))
print({k:j[k] for k in ('value','ms_per_step','launches_per_step')}, j['roofline']['achieved'], j['roofline']['frac'], j['step_roofline']['frac'], j['e2e']['value'])
PY
tail -3 gpurun_out/bench.err
