#!/bin/bash
# round 2, call P (8 GPUs): tensor-parallel parity at 4 and 8 ranks, strong-scaling bench lines at 8 and 4 GPUs (7B Q4_0 = the driver's SCALE path, 13B Q5_1 = BASELINE configs[3])
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( timeout 400 python -m pytest tests/test_gpu_tp.py -q -m gpu -p no:cacheprovider -x -k "4-mha or 8-mha" ) > gpurun_out/r02p_tp_tests.log 2>&1; echo "tp pytest rc=$?"; tail -4 gpurun_out/r02p_tp_tests.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29718 bench.py --gpus 8 > gpurun_out/r02p_bench_tp8_7b.json 2> gpurun_out/r02p_bench_tp8_7b.err; echo "tp8 7b rc=$?"; tail -2 gpurun_out/r02p_bench_tp8_7b.err | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29728 bench.py --gpus 8 --model 13b-q5_1 --layers 40 > gpurun_out/r02p_bench_tp8_13b.json 2> gpurun_out/r02p_bench_tp8_13b.err; echo "tp8 13b rc=$?"; tail -2 gpurun_out/r02p_bench_tp8_13b.err | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29714 bench.py --gpus 4 > gpurun_out/r02p_bench_tp4_7b.json 2> gpurun_out/r02p_bench_tp4_7b.err; echo "tp4 7b rc=$?"; tail -2 gpurun_out/r02p_bench_tp4_7b.err | cut -c1-200
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02p_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); x=d.get('exchange',{})
        print(f, round(d['value'],1), round(d['ms_per_step'],4), 'nowait', round(x.get('ms_per_step_without_tag_waits',0),3), 'local', round(x.get('ms_per_step_local_stores_only',0),3))
    except Exception as e: print(f, 'ERR', e)
PY
