"""Which side of the NeoX seam comparison is non-deterministic?  N iterations of: two CPU models and two seam models on the same 19-token prompt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import bindings as B
from oracle import synth
tag = sys.argv[1] if len(sys.argv) > 1 else "0"
ref, seam = B.RefLib("ref"), B.RefLib("seam")
bad = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 12):
    cfg = ("neox-tiny", "neox-tiny-seq")[it % 2]
    t = (B.Q4_0, B.Q4_1, B.Q5_0, B.Q8_0)[(it // 2) % 4]
    hp, tens = synth.make_neox(synth.NEOX_CONFIGS[cfg], t, ref.quantize)
    toks = np.random.default_rng(17).integers(0, hp["n_vocab"], 30, dtype=np.int32)
    outs = {}
    for name, lib, kw in (("c1", ref, {}), ("g1", seam, dict(use_gpu=1)), ("c2", ref, {}), ("g2", seam, dict(use_gpu=1))):
        m = lib.neox(hp, tens, n_threads=2, n_batch=32, **kw)
        outs[name] = m.eval(toks[:19]).copy()
        m.close()
    eq = lambda a, b: np.array_equal(outs[a].view(np.uint32), outs[b].view(np.uint32))
    line = f"[{tag}] it={it} {cfg} t={t}: c1==c2 {eq('c1','c2')}  g1==g2 {eq('g1','g2')}  c1==g1 {eq('c1','g1')}  c1==g2 {eq('c1','g2')}  nan(c1,c2,g1,g2)={[int(np.isnan(outs[k]).any()) for k in ('c1','c2','g1','g2')]}"
    if not (eq('c1', 'c2') and eq('g1', 'g2') and eq('c1', 'g1')):
        bad += 1
        print("MISMATCH " + line, flush=True)
    else:
        print(line, flush=True)
print(f"[{tag}] mismatching iterations: {bad}")
