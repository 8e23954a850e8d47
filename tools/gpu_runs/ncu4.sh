#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:mm_exact_mma -s 8 -c 4 -f -o gpurun_out/r01e_mm_exact_mma python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench6.log 2>&1
ls -la gpurun_out/r01e_mm_exact_mma.ncu-rep
