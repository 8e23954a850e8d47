#!/bin/bash
mkdir -p gpurun_out
# launch list of ~2 decode steps (skip synth + the KV-filling prefill + warm-up)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1700 -c 1100 --csv --log-file gpurun_out/r01_launches_decode.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prefill > gpurun_out/ncu_bench1.log 2>&1
# full capture of the exact streaming mat-vec (w13-sized and wo-sized launches) and the fast mat-vec
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mmv_exact_stream -s 10 -c 4 -o gpurun_out/r01_mmv_exact_stream python bench.py --layers 2 --steps 2 --warmup 3 --no-cpu-baseline --no-prefill > gpurun_out/ncu_bench2.log 2>&1
ls -la gpurun_out/ | tail -8
