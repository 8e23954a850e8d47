#!/bin/bash
# round 2, call I (2 GPUs): the same two-rank parity + bench with the exchange rewritten as tagged 8-byte units
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
nvidia-smi topo -m 2>/dev/null | head -6
( timeout 700 python -m pytest tests/test_gpu_tp.py -q -m gpu -p no:cacheprovider -x ) > gpurun_out/r02i_tp_tests.log 2>&1; echo "tp pytest rc=$?"; tail -25 gpurun_out/r02i_tp_tests.log | cut -c1-600
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/r02i_bench_tp2_7b.json 2> gpurun_out/r02i_bench_tp2_7b.err; echo "tp2 7b rc=$?"; tail -6 gpurun_out/r02i_bench_tp2_7b.err | cut -c1-300; head -c 900 gpurun_out/r02i_bench_tp2_7b.json; echo
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --model 13b-q5_1 --layers 40 > gpurun_out/r02i_bench_tp2_13b.json 2> gpurun_out/r02i_bench_tp2_13b.err; echo "tp2 13b rc=$?"; tail -6 gpurun_out/r02i_bench_tp2_13b.err | cut -c1-300; head -c 900 gpurun_out/r02i_bench_tp2_13b.json; echo
