#!/bin/bash
# compute-sanitizer over the NeoX graph through the seam (the flaky (0,19) mismatch of tests/test_seam_gpt2_neox.py)
mkdir -p gpurun_out
timeout 280 compute-sanitizer --tool initcheck --print-limit 3000 python tools/gpt2_debug.py neox > gpurun_out/san_init.log 2>&1
grep -E "     at |ERROR SUMMARY|equal|DIFF" gpurun_out/san_init.log | sort | uniq -c | sort -rn | head -20
timeout 280 compute-sanitizer --tool memcheck --print-limit 8 python tools/gpt2_debug.py neox > gpurun_out/san_mem.log 2>&1
grep -E "Invalid|at 0x|ERROR SUMMARY|equal|DIFF|in .*kernel|by thread" gpurun_out/san_mem.log | head -20
