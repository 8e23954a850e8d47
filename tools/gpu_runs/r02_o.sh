#!/bin/bash
# round 2, call O (2 GPUs): which kernels may skip the grid-completion wait (B200_TP_RELAX 0 / 1 / 3), 7B and 13B
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_tp.py -q -m gpu -p no:cacheprovider -x ) > gpurun_out/r02o_tp_tests.log 2>&1; echo "tp pytest rc=$?"; tail -3 gpurun_out/r02o_tp_tests.log | cut -c1-300
for r in 1 0 3; do
  B200_TP_RELAX=$r timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$r bench.py --gpus 2 > gpurun_out/r02o_bench_tp2_7b_relax$r.json 2> gpurun_out/r02o_bench_tp2_7b_relax$r.err; echo "tp2 7b relax=$r rc=$?"
done
for r in 1 0; do
  B200_TP_RELAX=$r timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$r bench.py --gpus 2 --model 13b-q5_1 --layers 40 > gpurun_out/r02o_bench_tp2_13b_relax$r.json 2> gpurun_out/r02o_bench_tp2_13b_relax$r.err; echo "tp2 13b relax=$r rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02o_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); x=d.get('exchange',{})
        print(f, round(d['value'],1), round(d['ms_per_step'],4), 'nowait', round(x.get('ms_per_step_without_tag_waits',0),3), 'local', round(x.get('ms_per_step_local_stores_only',0),3))
    except Exception as e: print(f, 'ERR', e)
PY
