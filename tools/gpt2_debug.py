import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from anywhere: the repo root holds llm_b200/ and oracle/
import sys
import numpy as np
from oracle import bindings as B
from oracle import synth
which = sys.argv[1] if len(sys.argv) > 1 else "gpt2"
ref, seam = B.RefLib("ref"), B.RefLib("seam")
if which == "gpt2":
    hp, tens = synth.make_gpt2(synth.GPT2_CONFIGS["gpt2-tiny"], B.Q4_0, ref.quantize)
    mk = lambda lib, **kw: lib.gpt2(hp, tens, **kw)
else:
    hp, tens = synth.make_neox(synth.NEOX_CONFIGS["neox-tiny"], B.Q4_0, ref.quantize)
    mk = lambda lib, **kw: lib.neox(hp, tens, **kw)
toks = np.random.default_rng(11).integers(0, hp["n_vocab"], 30, dtype=np.int32)
mc = mk(ref, n_threads=2, n_batch=32)
mg = mk(seam, use_gpu=1, n_threads=2, n_batch=32)
for lo, hi in ((0, 21), (21, 22), (22, 23), (23, 30)):
    want, got = mc.eval(toks[lo:hi]), mg.eval(toks[lo:hi])
    print(which, lo, hi, "equal" if np.array_equal(got.view(np.uint32), want.view(np.uint32)) else "DIFF %g" % np.abs(got - want).max(), flush=True)
