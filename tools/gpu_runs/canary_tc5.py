"""one small tcgen05 GEMM launch (exits non-zero on a barrier timeout): run under `timeout 60` before anything that would launch it thousands of times"""
import ctypes as C, sys
sys.path.insert(0, ".")
from llm_b200 import _lib
L = _lib.lib()
L.b200_op_bench_mul_mat.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_float)]
L.b200_init(0)
ms = C.c_float()
for wtype in (2, 7):
    rc = L.b200_op_bench_mul_mat(wtype, 1024, 256, 256, 7, 1, C.byref(ms))
    print("canary", wtype, rc, ms.value, flush=True)
    if rc != 0:
        sys.exit(3)
