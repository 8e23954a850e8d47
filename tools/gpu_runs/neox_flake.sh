#!/bin/bash
mkdir -p gpurun_out
for k in 1 2 3 4; do (timeout 400 python tools/neox_stress.py $k 16 > gpurun_out/stress_$k.log 2>&1 &) ; done; sleep 80
grep -h "MISMATCH\|mismatching" gpurun_out/stress_*.log | head -12
for i in 1 2 3 4; do
  timeout 300 python -m pytest tests/test_seam_gpt2_neox.py -m gpu -q -n 4 -p no:cacheprovider 2>&1 | grep -E "passed|failed|AssertionError" | head -4
done
( timeout 1200 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider ) 2>&1 | grep -E "passed|failed|AssertionError" | head -4
