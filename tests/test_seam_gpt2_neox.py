"""GPT-2 through the reference (SURVEY.md §8 a26 / BASELINE.json configs[0]) -- oracle/ref_gpt2.c restates crates/models/gpt2 on the ggml C API.

CPU: configs[0] itself -- GPT-2 117M geometry, Q4_0, a 32-token prompt and one decode step on the reference ggml CPU path ("plumbing": the
graph must build and run under ggml's own asserts; logits finite; batched and token-by-token evaluation agree to the reference's own tolerance).
GPU: the same graph executed by the reference's ggml.c over OUR ggml_cuda_* seam (use_gpu) against the CPU build, bit for bit.
"""
import numpy as np
import pytest

from oracle import bindings as B
from oracle import synth

pytestmark = pytest.mark.skipif(not B.have_ref("ref"), reason="oracle/_ref not built")


def test_config0_gpt2_117m_q4_0_prompt32_cpu(ref):
    hp, tens = synth.make_gpt2(synth.GPT2_CONFIGS["gpt2-117m"], B.Q4_0, ref.quantize)
    toks = np.random.default_rng(0x70CE11).integers(0, hp["n_vocab"], 33, dtype=np.int32)
    m = ref.gpt2(hp, tens, n_threads=8, n_batch=32)
    logits = m.eval(toks[:32])
    assert logits.shape == (32, hp["n_vocab"]) and np.isfinite(logits).all()
    one = m.eval(toks[32:33])
    assert one.shape == (1, hp["n_vocab"]) and np.isfinite(one).all()
    assert float(np.abs(logits).max()) > 1e-3                      # a real forward pass, not zeros
    m.close()


def test_gpt2_batched_equals_incremental_cpu(ref):
    """prefill of 9 tokens == 9 single-token evaluations on the reference itself (same kernels, same order per row)"""
    hp, tens = synth.make_gpt2(synth.GPT2_CONFIGS["gpt2-tiny"], B.Q4_0, ref.quantize, lm_head=True)
    toks = np.random.default_rng(5).integers(0, hp["n_vocab"], 9, dtype=np.int32)
    a = ref.gpt2(hp, tens, n_threads=2, n_batch=16)
    b = ref.gpt2(hp, tens, n_threads=2, n_batch=16)
    batched = a.eval(toks)
    inc = np.concatenate([b.eval(toks[i:i + 1]) for i in range(9)])
    assert np.allclose(batched, inc, rtol=0, atol=5e-2 * np.abs(batched).max())       # the chaos bound (tests/test_chaos.py), not bit-exact: mul_mat
    a.close(); b.close()                                                              # quantizes activations per row either way, but KQ sizes differ


@pytest.mark.gpu
@pytest.mark.parametrize("name,lm_head", [("q4_0", False), ("q5_1", True), ("q8_0", False)])
def test_gpt2_reference_executor_over_our_seam(name, lm_head):
    """Gpt2::evaluate with use_gpu: the reference's graph executor, every offloaded node in our kernels (LayerNorm, gelu, biases, f16 V transpose copy).
    The embedding tables stay on the host (see oracle/ref_gpt2.c: as written, the reference aborts in ggml.c:14589 because no CUDA get_rows exists)."""
    t = B.QUANT_TYPES[name]
    ref, seam = B.RefLib("ref"), B.RefLib("seam")
    hp, tens = synth.make_gpt2(synth.GPT2_CONFIGS["gpt2-tiny"], t, ref.quantize, lm_head=lm_head)
    toks = np.random.default_rng(11).integers(0, hp["n_vocab"], 30, dtype=np.int32)
    mc = ref.gpt2(hp, tens, n_threads=2, n_batch=32)
    mg = seam.gpt2(hp, tens, use_gpu=1, n_threads=2, n_batch=32)
    for lo, hi in ((0, 21), (21, 22), (22, 23), (23, 30)):
        want, got = mc.eval(toks[lo:hi]), mg.eval(toks[lo:hi])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, lo, hi, float(np.abs(got - want).max()))
    mg.close(); mc.close()


def test_neox_graph_runs_on_the_reference_cpu(ref):
    """GptNeoX::evaluate restated (parallel and sequential residual, rope mode 2 on n_rot < head size): builds and runs under ggml's asserts"""
    for cfg in ("neox-tiny", "neox-tiny-seq"):
        hp, tens = synth.make_neox(synth.NEOX_CONFIGS[cfg], B.Q5_0, ref.quantize)
        toks = np.random.default_rng(3).integers(0, hp["n_vocab"], 12, dtype=np.int32)
        m = ref.neox(hp, tens, n_threads=2, n_batch=16)
        a = m.eval(toks[:11]); b = m.eval(toks[11:12])
        assert a.shape == (11, hp["n_vocab"]) and np.isfinite(a).all() and np.isfinite(b).all() and float(np.abs(a).max()) > 1e-3
        m.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["q4_0", "q4_1", "q5_0", "q5_1", "q8_0"])
@pytest.mark.parametrize("cfg", ["neox-tiny", "neox-tiny-seq"])
def test_neox_quant_format_sweep_over_our_seam(cfg, name):
    """BASELINE.json configs[4] (the five block formats on GPT-NeoX) at test size: the reference executor over our seam vs its CPU build, bit for bit."""
    t = B.QUANT_TYPES[name]
    ref, seam = B.RefLib("ref"), B.RefLib("seam")
    hp, tens = synth.make_neox(synth.NEOX_CONFIGS[cfg], t, ref.quantize)
    toks = np.random.default_rng(17).integers(0, hp["n_vocab"], 30, dtype=np.int32)
    mc = ref.neox(hp, tens, n_threads=2, n_batch=32)
    mg = seam.neox(hp, tens, use_gpu=1, n_threads=2, n_batch=32)
    for lo, hi in ((0, 19), (19, 20), (20, 21), (21, 30)):
        want, got = mc.eval(toks[lo:hi]), mg.eval(toks[lo:hi])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (cfg, name, lo, hi, float(np.abs(got - want).max()))
    mg.close(); mc.close()
