#!/bin/bash
# round 2, call B: ncu source-level profile of the tcgen05 exact GEMM (qkv shape), new published-config parity tests
mkdir -p gpurun_out
cat > /tmp/one_gemm.py <<'PY'
import ctypes as C, sys
sys.path.insert(0, ".")
from llm_b200 import _lib
L = _lib.lib()
L.b200_op_bench_mul_mat.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
L.b200_init(0)
ms = C.c_float()
print(L.b200_op_bench_mul_mat(2, 4096, 12288, 512, 7, 2, C.byref(ms)), ms.value)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mm_exact_tc5 -s 2 -c 1 -o gpurun_out/r02b_tc5 python /tmp/one_gemm.py > gpurun_out/r02b_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02b_ncu.log
( timeout 1500 python -m pytest tests/test_gpu_llama.py -q -m gpu -p no:cacheprovider -k "published or bucket or rope_overrides or 13b" -x ) > gpurun_out/r02b_tests.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02b_tests.log
