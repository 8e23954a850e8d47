#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu.log
tail -5 gpurun_out/smoke.log; grep -E "AssertionError: \(|passed|failed|FAILED" gpurun_out/test_gpu.log | head -60
