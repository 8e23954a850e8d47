#!/bin/bash
# round 2, call N (1 GPU): bisect of the 2 % single-GPU decode regression (the same bench from the trees of five commits on one box); K-quant parity tests
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kquants.py -q -m gpu -p no:cacheprovider -x ) > gpurun_out/r02n_kquant_tests.log 2>&1; echo "kquant pytest rc=$?"; tail -6 gpurun_out/r02n_kquant_tests.log | cut -c1-400
for c in r01 03c972c 81b1914 502418f ac36e14; do
  if [ -d tools/ab/$c ]; then
    extra="--no-prefill"; grep -q -- "--no-prefill" tools/ab/$c/bench.py || extra=""
    ( cd tools/ab/$c && timeout 300 python bench.py --no-cpu-baseline $extra > ../../../gpurun_out/r02n_decode_$c.json 2> ../../../gpurun_out/r02n_decode_$c.err )
    echo "$c: $(grep -h 'decode@1' gpurun_out/r02n_decode_$c.err | tail -1)"
  fi
done
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02n_decode_head.json 2> gpurun_out/r02n_decode_head.err; echo "HEAD: $(grep -h 'decode@1' gpurun_out/r02n_decode_head.err | tail -1)"
( cd tools/ab/r01 && timeout 300 python bench.py --no-cpu-baseline --no-prefill > ../../../gpurun_out/r02n_decode_r01_again.json 2> ../../../gpurun_out/r02n_decode_r01_again.err ); echo "r01 again: $(grep -h 'decode@1' gpurun_out/r02n_decode_r01_again.err | tail -1)"
