#!/bin/bash
# round 2, call C: tcgen05 exact GEMM v2 (pipelined epilogue): parity, timings; new bench.py decode + prefill lines
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "tcgen05" -x ) > gpurun_out/r02c_tc5_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02c_tc5_tests.log
timeout 300 python tools/prefill_gemm_bench.py 2 512 > gpurun_out/r02c_gemm_bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/r02c_gemm_bench.log
timeout 300 python tools/prefill_gemm_bench.py 7 512 7 > gpurun_out/r02c_gemm_bench_q51.log 2>&1; tail -3 gpurun_out/r02c_gemm_bench_q51.log
( time timeout 900 python bench.py > gpurun_out/r02c_bench_decode.json 2> gpurun_out/r02c_bench_decode.err ) 2>&1 | grep real; tail -25 gpurun_out/r02c_bench_decode.err; head -c 1500 gpurun_out/r02c_bench_decode.json; echo
