#!/bin/bash
# round 2, call M (2 GPUs): decode consumer without I2F (seed-magic FADD): parity + bench; timeline A/B vs the round-1 tree; TP with sampled back-off polling
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_llama.py tests/test_gpu_neox.py tests/test_gpu_ops.py -q -m gpu -p no:cacheprovider -x -k "not tcgen05" ) > gpurun_out/r02m_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02m_tests.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02m_bench_decode.json 2> gpurun_out/r02m_bench_decode.err; tail -1 gpurun_out/r02m_bench_decode.err
timeout 200 python tools/decode_timeline.py > gpurun_out/r02m_timeline.txt 2>&1; head -12 gpurun_out/r02m_timeline.txt
if [ -d tools/ab/r01 ]; then ( cd tools/ab/r01 && timeout 200 python tools/decode_timeline.py > ../../../gpurun_out/r02m_timeline_r01tree.txt 2>&1 ); head -12 gpurun_out/r02m_timeline_r01tree.txt; fi
( timeout 600 python -m pytest tests/test_gpu_tp.py -q -m gpu -p no:cacheprovider -x ) > gpurun_out/r02m_tp_tests.log 2>&1; echo "tp pytest rc=$?"; tail -5 gpurun_out/r02m_tp_tests.log | cut -c1-400
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 > gpurun_out/r02m_bench_tp2_7b.json 2> gpurun_out/r02m_bench_tp2_7b.err; echo "tp2 7b rc=$?"; tail -3 gpurun_out/r02m_bench_tp2_7b.err | cut -c1-300
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --model 13b-q5_1 --layers 40 > gpurun_out/r02m_bench_tp2_13b.json 2> gpurun_out/r02m_bench_tp2_13b.err; echo "tp2 13b rc=$?"; tail -3 gpurun_out/r02m_bench_tp2_13b.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02m_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value'],1), round(d['ms_per_step'],4), d.get('exchange'))
    except Exception as e: print(f, 'ERR', e)
PY
