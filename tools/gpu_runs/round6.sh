#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 --timeout 600 -p no:cacheprovider -k "decode_kernel or native_vs_oracle or synthesized or session_semantics or golden" > gpurun_out/test_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu.log
grep -E "AssertionError|passed|failed|FAILED|Error|error" gpurun_out/test_gpu.log | head -20
python tools/decode_prof.py 4 2>&1 | tail -22
