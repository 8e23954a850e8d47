#!/bin/bash
# round 2, call K (2 GPUs): decode kernels templated on TP (single-GPU graph = round-1 code), TP consumers with batched tag polling
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02k_bench_decode.json 2> gpurun_out/r02k_bench_decode.err; tail -2 gpurun_out/r02k_bench_decode.err
( timeout 600 python -m pytest tests/test_gpu_tp.py -q -m gpu -p no:cacheprovider -x ) > gpurun_out/r02k_tp_tests.log 2>&1; echo "tp pytest rc=$?"; tail -5 gpurun_out/r02k_tp_tests.log | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/r02k_bench_tp2_7b.json 2> gpurun_out/r02k_bench_tp2_7b.err; echo "tp2 7b rc=$?"; tail -4 gpurun_out/r02k_bench_tp2_7b.err | cut -c1-300; head -c 700 gpurun_out/r02k_bench_tp2_7b.json; echo
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --model 13b-q5_1 --layers 40 > gpurun_out/r02k_bench_tp2_13b.json 2> gpurun_out/r02k_bench_tp2_13b.err; echo "tp2 13b rc=$?"; tail -4 gpurun_out/r02k_bench_tp2_13b.err | cut -c1-300; head -c 700 gpurun_out/r02k_bench_tp2_13b.json; echo
