#!/bin/bash
# round 2, call A: tcgen05 plumbing probe, parity of the tcgen05 exact GEMM, kernel-only prefill GEMM timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
timeout 120 ./tools/tc5_probe > gpurun_out/r02a_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02a_probe.log
( timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "tcgen05" -x ) > gpurun_out/r02a_tc5_tests.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02a_tc5_tests.log
timeout 300 python tools/prefill_gemm_bench.py 2 512 > gpurun_out/r02a_gemm_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/r02a_gemm_bench.log
