#!/bin/bash
# first GPU contact: environment, smoke, parity tests, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/env.txt 2>&1
nproc >> gpurun_out/env.txt; lscpu | grep -E "Model name|Socket|Core|Thread" >> gpurun_out/env.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q -n 4 --timeout 900 -p no:cacheprovider > gpurun_out/test_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/test_gpu.log
timeout 900 python bench.py --steps 32 --warmup 4 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/smoke.log; tail -30 gpurun_out/test_gpu.log; cat gpurun_out/bench.json; tail -15 gpurun_out/bench.err
