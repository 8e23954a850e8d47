#!/bin/bash
mkdir -p gpurun_out
python tools/decode_prof.py 4 2>&1 | tail -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:llama_decode -s 3 -c 1 -o gpurun_out/r01b_decode_kernel python tools/decode_prof.py 2 > gpurun_out/ncu_decode.log 2>&1
ls -la gpurun_out/*.ncu-rep
