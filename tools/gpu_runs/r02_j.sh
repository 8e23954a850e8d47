#!/bin/bash
# round 2, call J (1 GPU): exact tcgen05 GEMM v3b (barrier probe hoisted), decode with TpCtx in constant memory
mkdir -p gpurun_out
timeout 90 python tools/gpu_runs/canary_tc5.py > gpurun_out/r02j_canary.log 2>&1 || { echo "CANARY FAILED"; cat gpurun_out/r02j_canary.log; exit 1; }
( timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "tcgen05" ) > gpurun_out/r02j_tc5_tests.log 2>&1; echo "tc5 pytest rc=$?"; tail -4 gpurun_out/r02j_tc5_tests.log
timeout 200 python tools/prefill_gemm_bench.py 2 512 7 > gpurun_out/r02j_gemm_bench.log 2>&1; cat gpurun_out/r02j_gemm_bench.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/r02j_bench_decode.json 2> gpurun_out/r02j_bench_decode.err; tail -3 gpurun_out/r02j_bench_decode.err; head -c 600 gpurun_out/r02j_bench_decode.json; echo
timeout 400 python bench.py --metric prefill --no-cpu-baseline > gpurun_out/r02j_bench_prefill.json 2> gpurun_out/r02j_bench_prefill.err; tail -3 gpurun_out/r02j_bench_prefill.err; head -c 600 gpurun_out/r02j_bench_prefill.json; echo
