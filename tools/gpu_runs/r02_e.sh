#!/bin/bash
# round 2, call E: ncu of the tcgen05 exact GEMM v2, fast tcgen05 GEMM (canary + tests + timings), weight quantizer, then the whole GPU suite
mkdir -p gpurun_out
timeout 90 python tools/gpu_runs/canary_tc5.py > gpurun_out/r02e_canary.log 2>&1 || { echo "CANARY FAILED"; cat gpurun_out/r02e_canary.log; exit 1; }
cat > /tmp/one_gemm.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, ".")
from llm_b200 import _lib
L = _lib.lib()
L.b200_op_bench_mul_mat.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
L.b200_init(0)
ms = C.c_float()
impl = int(sys.argv[1])
print(L.b200_op_bench_mul_mat(2, 4096, 12288, 512, impl, 2, C.byref(ms)), ms.value)
PY
timeout 60 python /tmp/one_gemm.py 8 > gpurun_out/r02e_fast_canary.log 2>&1; echo "fast canary rc=$?"; cat gpurun_out/r02e_fast_canary.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mm_exact_tc5 -s 2 -c 1 -o gpurun_out/r02e_tc5v2 python /tmp/one_gemm.py 7 > gpurun_out/r02e_ncu.log 2>&1; echo "ncu rc=$?"
( timeout 300 python -m pytest tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -k "fast_tcgen05 or weight_quantizer" ) > gpurun_out/r02e_fast_tests.log 2>&1; echo "fast/quantizer pytest rc=$?"; tail -6 gpurun_out/r02e_fast_tests.log
timeout 200 python tools/prefill_gemm_bench.py 2 512 7,8 > gpurun_out/r02e_gemm_bench.log 2>&1; cat gpurun_out/r02e_gemm_bench.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mm_fast_tc5 -s 2 -c 1 -o gpurun_out/r02e_fast python /tmp/one_gemm.py 8 > gpurun_out/r02e_ncu_fast.log 2>&1; echo "ncu fast rc=$?"
( time timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider ) > gpurun_out/r02e_test_gpu_all.log 2>&1; echo "full pytest rc=$?"; tail -15 gpurun_out/r02e_test_gpu_all.log
