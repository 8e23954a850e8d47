#!/bin/bash
# round 2, call Q (1 GPU): driver-like validation (smoke, the whole GPU suite, default bench lines) + the ncu launch lists of the bench commands + v3b GEMM capture  [NOT RUN: the GPU budget of the round ended with call P; kept as the recipe for the next round]
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02q_smoke.log
( timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r02q_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r02q_gpu_suite.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r02q_bench_decode.json 2> gpurun_out/r02q_bench_decode.err; echo "bench decode rc=$?"; tail -2 gpurun_out/r02q_bench_decode.err | cut -c1-300
timeout 600 python bench.py --metric prefill --no-cpu-baseline > gpurun_out/r02q_bench_prefill.json 2> gpurun_out/r02q_bench_prefill.err; echo "bench prefill rc=$?"; tail -1 gpurun_out/r02q_bench_prefill.err | cut -c1-300
B200_PREFILL_GEMM=mma timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02q_bench_decode_mma_prefill.json 2> gpurun_out/r02q_bench_decode_mma_prefill.err; echo "decode after mma prefill: $(grep -h 'decode@1' gpurun_out/r02q_bench_decode_mma_prefill.err | tail -1)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02q_launches_decode.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_ncu_decode.log 2>&1; echo "ncu decode rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02q_launches_prefill.csv python bench.py --metric prefill --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02q_ncu_prefill.log 2>&1; echo "ncu prefill rc=$?"
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
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mm_exact_tc5 -s 2 -c 1 -o gpurun_out/r02q_mm_exact_tc5_v3b python /tmp/one_gemm.py 7 > gpurun_out/r02q_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
wc -l gpurun_out/r02q_launches_*.csv
