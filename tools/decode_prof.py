import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from anywhere: the repo root holds llm_b200/ and oracle/
import ctypes as C, os, sys
import numpy as np
os.environ["B200_DECODE_PROF"] = "1"
import llm_b200
from llm_b200 import _lib
L = _lib.lib()
hp = dict(n_vocab=32000, n_embd=4096, n_head=32, n_head_kv=32, n_layer=int(sys.argv[1]) if len(sys.argv) > 1 else 4, n_rot=128, n_ff=11008, wtype=2)
m = llm_b200.Llama(hp, llm_b200.ModelParameters(context_size=2048)); m.synthesize(1)
s = m.start_session(llm_b200.InferenceSessionConfig(n_batch=512))
toks = np.random.default_rng(0).integers(0, 32000, 513, dtype=np.int32)
s.evaluate(toks[:512])
for _ in range(5):
    s.rewind(512); s.evaluate(toks[512:513])
buf = (C.c_ulonglong * 128)()
L.b200_session_decode_profile(s._s, buf)
t = np.array(buf[:], dtype=np.float64)
t0 = t[0]
names = ["A qkv", "B kq", "C softmax+kqv", "D wo", "E w13", "F w2"]
has_pro = [True, False, False, True, True, True]
print("embed+barrier %.1f us" % ((t[1] - t0) / 1e3))
idx = 2; prev = t[1]
tot = {}
for il in range(min(hp["n_layer"], 3)):
    for ph in range(6):
        pro = 0.0
        if has_pro[ph]:
            pro = (t[idx] - prev) / 1e3; prev = t[idx]; idx += 1
        work, bar = (t[idx] - prev) / 1e3, (t[idx + 1] - t[idx]) / 1e3
        print(f"layer {il} {names[ph]:14s} prologue {pro:6.1f} us  work {work:7.1f} us   barrier wait {bar:6.1f} us")
        prev = t[idx + 1]; idx += 2
print("token total %.1f us" % ((t[127] - t0) / 1e3))
