"""GPU debugging aid: per-stage comparison of the native session against the oracle on one tiny model."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from anywhere: the repo root holds llm_b200/ and oracle/

import ctypes as C
import sys
import numpy as np
import llm_b200
from oracle import bindings as B, synth

orc = B.Oracle()
cfg = sys.argv[1] if len(sys.argv) > 1 else "tiny"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 33
hp, tens = synth.make_llama(synth.CONFIGS[cfg], B.Q4_0, orc.quantize)
toks = synth.make_tokens(hp, n + 8)
e, f, nh = hp["n_embd"], hp["n_ff"], hp["n_head"]
gqa = e // (hp["n_head"] // hp["n_head_kv"])
m = llm_b200.Llama(hp, llm_b200.ModelParameters(context_size=hp["n_ctx"]), tens)
s = m.start_session(llm_b200.InferenceSessionConfig(n_batch=256))
L = m.L


def run(chunks, il):
    for stage in range(1, 12):
        mo = orc.llama(hp, tens)
        s.rewind(0)
        past = 0
        for ci, ch in enumerate(chunks):
            last = ci == len(chunks) - 1
            nn, n_kv = ch.size, past + ch.size
            counts = {1: nn * e, 2: nn * (e + 2 * gqa), 3: nn * (e + 2 * gqa), 4: nh * nn * n_kv, 5: nh * nn * n_kv, 6: nn * e, 7: nn * e, 8: nn * e,
                      9: nn * 2 * f, 10: nn * f, 11: nn * e}
            if last:
                L.b200_session_set_tap(s._s, il, stage)
                g_logits = s.evaluate(ch, all_logits=True)
                buf = np.zeros(counts[stage], np.float32)
                got = L.b200_session_read_tap(s._s, buf.ctypes.data, buf.size)
                L.b200_session_set_tap(s._s, -2, 0)
                ocount = counts[stage] if stage != 3 else nn * (e + gqa)
                c_logits, tap = mo.eval_tap(ch, il, stage, ocount)
                if stage == 2:
                    g = buf.reshape(nn, e + 2 * gqa); g = np.concatenate([g[:, :e].ravel(), g[:, e:e + gqa].ravel(), g[:, e + gqa:].ravel()])
                elif stage == 3:
                    g = buf.reshape(nn, e + 2 * gqa); g = np.concatenate([g[:, :e].ravel(), g[:, e:e + gqa].ravel()])
                elif stage == 9:
                    g = buf.reshape(nn, 2 * f); g = np.concatenate([g[:, :f].ravel(), g[:, f:].ravel()])
                else:
                    g = buf
                fin = np.isfinite(tap)
                err = np.abs(g[fin] - tap[fin]).max() / max(np.abs(tap[fin]).max(), 1e-30)
                nz = int((g[fin] != tap[fin]).sum())
                print(f"layer {il} stage {stage:2d}: got {got} floats, max-rel err {err:.3e}, differing {nz}/{tap.size}, logits err {np.abs(g_logits - c_logits).max() / np.abs(c_logits).max():.3e}")
            else:
                s.evaluate(ch)
                mo.eval(ch)
            past += ch.size


print("== single prefill of", n)
run([toks[:n]], 0)
print("== chunked: 16 then 8")
run([toks[:16], toks[16:24]], 0)
print("== decode after 20")
run([toks[:20], toks[20:21]], 0)
