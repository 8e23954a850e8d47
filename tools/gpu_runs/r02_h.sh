#!/bin/bash
# round 2, call H (1 GPU): exact tcgen05 GEMM v4 (24-row tiles, 12 epilogue warps): canary, parity, timings, ncu, prefill bench line; decode regression check
mkdir -p gpurun_out
timeout 90 python tools/gpu_runs/canary_tc5.py > gpurun_out/r02h_canary.log 2>&1 || { echo "CANARY FAILED"; cat gpurun_out/r02h_canary.log; exit 1; }
cat > /tmp/one_gemm.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, ".")
from llm_b200 import _lib
L = _lib.lib()
L.b200_op_bench_mul_mat.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
L.b200_init(0)
ms = C.c_float()
print(L.b200_op_bench_mul_mat(2, 4096, 12288, 512, int(sys.argv[1]), 2, C.byref(ms)), ms.value)
PY
( timeout 400 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "tcgen05" ) > gpurun_out/r02h_tc5_tests.log 2>&1; echo "tc5 pytest rc=$?"; tail -4 gpurun_out/r02h_tc5_tests.log
timeout 200 python tools/prefill_gemm_bench.py 2 512 7 > gpurun_out/r02h_gemm_bench.log 2>&1; cat gpurun_out/r02h_gemm_bench.log
timeout 100 python tools/prefill_gemm_bench.py 7 512 7 > gpurun_out/r02h_gemm_bench_q51.log 2>&1; tail -1 gpurun_out/r02h_gemm_bench_q51.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mm_exact_tc5 -s 2 -c 1 -o gpurun_out/r02h_tc5v4 python /tmp/one_gemm.py 7 > gpurun_out/r02h_ncu.log 2>&1; echo "ncu rc=$?"
( timeout 600 python -m pytest tests/test_gpu_llama.py tests/test_gpu_neox.py -q -m gpu -p no:cacheprovider -x ) > gpurun_out/r02h_model_tests.log 2>&1; echo "model pytest rc=$?"; tail -4 gpurun_out/r02h_model_tests.log
timeout 400 python bench.py --metric prefill --no-cpu-baseline > gpurun_out/r02h_bench_prefill.json 2> gpurun_out/r02h_bench_prefill.err; tail -3 gpurun_out/r02h_bench_prefill.err
