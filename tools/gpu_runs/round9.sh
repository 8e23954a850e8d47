#!/bin/bash
mkdir -p gpurun_out
# full GPU test-suite (timed), then the full bench incl. cpu baseline, then ncu evidence of the graph decode path
( time timeout 1500 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider --durations=8 ) > gpurun_out/test_gpu_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu_full.log
grep -E "passed|failed|FAILED|real" gpurun_out/test_gpu_full.log | head
( time timeout 900 python bench.py ) > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?" >> gpurun_out/bench_full.err
cat gpurun_out/bench_full.json; tail -12 gpurun_out/bench_full.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 600 --csv --log-file gpurun_out/r01c_launches_decode_graph.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prefill > gpurun_out/ncu_bench3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mmv_fused -s 20 -c 4 -o gpurun_out/r01c_mmv_fused python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu-baseline --no-prefill > gpurun_out/ncu_bench4.log 2>&1
ls -la gpurun_out | tail -6
