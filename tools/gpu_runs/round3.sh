#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider -k "stream or session_semantics or native_vs_oracle or seam or golden" > gpurun_out/test_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu.log
grep -E "AssertionError|passed|failed|FAILED|Error" gpurun_out/test_gpu.log | head -40
timeout 900 python bench.py --steps 32 --warmup 4 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
