#!/bin/bash
# round 2, call D: tcgen05 GEMM v2 (canary first), NeoX native parity, new bench lines.  Every step has its own short timeout.
mkdir -p gpurun_out
timeout 90 python tools/gpu_runs/canary_tc5.py > gpurun_out/r02d_canary.log 2>&1; rc=$?; cat gpurun_out/r02d_canary.log
if [ $rc -ne 0 ]; then echo "CANARY FAILED rc=$rc: skipping the tcgen05 steps"; export B200_PREFILL_GEMM=mma; SKIP_TC5=1; fi
if [ -z "$SKIP_TC5" ]; then
  ( timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "tcgen05" -x ) > gpurun_out/r02d_tc5_tests.log 2>&1; echo "tc5 pytest rc=$?"; tail -4 gpurun_out/r02d_tc5_tests.log
  timeout 150 python tools/prefill_gemm_bench.py 2 512 > gpurun_out/r02d_gemm_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/r02d_gemm_bench.log
  timeout 100 python tools/prefill_gemm_bench.py 7 512 7 > gpurun_out/r02d_gemm_bench_q51.log 2>&1; tail -2 gpurun_out/r02d_gemm_bench_q51.log
fi
( timeout 600 python -m pytest tests/test_gpu_neox.py -q -m gpu -p no:cacheprovider -x ) > gpurun_out/r02d_neox_tests.log 2>&1; echo "neox pytest rc=$?"; tail -12 gpurun_out/r02d_neox_tests.log
( time timeout 500 python bench.py > gpurun_out/r02d_bench_decode.json 2> gpurun_out/r02d_bench_decode.err ) 2>&1 | grep real; tail -12 gpurun_out/r02d_bench_decode.err; head -c 1200 gpurun_out/r02d_bench_decode.json; echo
( time timeout 400 python bench.py --metric prefill --no-cpu-baseline > gpurun_out/r02d_bench_prefill.json 2> gpurun_out/r02d_bench_prefill.err ) 2>&1 | grep real; tail -5 gpurun_out/r02d_bench_prefill.err; head -c 1500 gpurun_out/r02d_bench_prefill.json; echo
