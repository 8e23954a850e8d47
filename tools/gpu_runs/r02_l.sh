#!/bin/bash
# round 2, call L (2 GPUs): A/B of the single-GPU decode against the round-1 tree on the same box; TP with relaxed grid waits (B200_TP_RELAX) on/off
mkdir -p gpurun_out
if [ -d tools/ab/r01 ]; then ( cd tools/ab/r01 && timeout 300 python bench.py --no-cpu-baseline --no-prefill > ../../../gpurun_out/r02l_bench_decode_r01tree.json 2> ../../../gpurun_out/r02l_bench_decode_r01tree.err ); tail -1 gpurun_out/r02l_bench_decode_r01tree.err; fi
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02l_bench_decode.json 2> gpurun_out/r02l_bench_decode.err; tail -1 gpurun_out/r02l_bench_decode.err
( timeout 600 python -m pytest tests/test_gpu_tp.py -q -m gpu -p no:cacheprovider -x ) > gpurun_out/r02l_tp_tests.log 2>&1; echo "tp pytest rc=$?"; tail -5 gpurun_out/r02l_tp_tests.log | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/r02l_bench_tp2_7b.json 2> gpurun_out/r02l_bench_tp2_7b.err; echo "tp2 7b rc=$?"; tail -3 gpurun_out/r02l_bench_tp2_7b.err | cut -c1-300
B200_TP_RELAX=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 > gpurun_out/r02l_bench_tp2_7b_norelax.json 2> gpurun_out/r02l_bench_tp2_7b_norelax.err; echo "tp2 7b norelax rc=$?"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --model 13b-q5_1 --layers 40 > gpurun_out/r02l_bench_tp2_13b.json 2> gpurun_out/r02l_bench_tp2_13b.err; echo "tp2 13b rc=$?"; tail -3 gpurun_out/r02l_bench_tp2_13b.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02l_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), d.get('exchange'))
    except Exception as e: print(f, 'ERR', e)
PY
