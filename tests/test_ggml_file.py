"""GGML / GGJT container: the parser and writer run on the host, so most of this file needs no GPU.

What is checked against the reference's loader rules (crates/ggml/src/format/loader.rs:160-281, crates/llm-base/src/loader.rs:459-484):
container magic + version, the 7 LLaMA hyperparameters, vocabulary with scores, tensor headers (dims ne0 first), 32-byte alignment of GGJT
tensor data, the Q4 `ne0 % 64` invariant, n_dims <= 2, unknown element types, the quantization-version rule, truncated files.
"""
import os
import struct

import numpy as np
import pytest

from oracle import bindings as B
from oracle import synth


def _model(orc, wtype=B.Q4_0, cfg="tiny"):
    hp, tens = synth.make_llama(synth.CONFIGS[cfg], wtype, orc.quantize)
    return hp, tens, synth.tensor_shapes(hp)


def _vocab(n):
    return [((b"tok%d" % i) if i % 7 else b"", float(-i) / 3.0) for i in range(n)]


def test_write_then_parse_roundtrip(orc, tmp_path):
    from llm_b200 import loader
    hp, tens, shapes = _model(orc)
    path = str(tmp_path / "tiny.ggjt")
    vocab = _vocab(hp["n_vocab"])
    loader.write_llama(path, hp, tens, shapes, vocabulary=vocab)
    f = loader.GgmlFile(path)
    assert f.container == ("ggjt", 3)
    got = f.llama_hyperparameters()
    for k in ("n_vocab", "n_embd", "n_head", "n_head_kv", "n_layer", "n_rot", "n_ff", "wtype"):
        assert got[k] == hp[k], k
    assert got["quantization_version"] == 2 and got["llama_ftype"] == 2 and got["n_mult"] == 256
    assert f.vocabulary() == [(t, np.float32(s)) for t, s in vocab]
    table = f.tensors()
    assert [t["name"] for t in table] == list(tens.keys())                       # file order = writer order
    for i, t in enumerate(table):
        shp = shapes[t["name"]]
        assert t["offset"] % 32 == 0                                             # loader.rs:259-263
        assert t["n_dims"] == len(shp) and t["ne"][0] == shp[-1] and t["ne"][1] == (shp[0] if len(shp) == 2 else 1)
        assert t["type"] == (0 if len(shp) == 1 else hp["wtype"])
        want = np.ascontiguousarray(tens[t["name"]]).view(np.uint8).ravel()
        assert t["nbytes"] == want.size and np.array_equal(f.tensor_bytes(i), want)
    f.close()


def _raw_file(path, magic=0x67676a74, version=3, hp=(4, 64, 256, 2, 1, 32, 2002), tokens=None, tensors=()):
    """hand-assembled file: tensors = [(n_dims, name, ftype, dims, payload_bytes)]"""
    with open(path, "wb") as fp:
        fp.write(struct.pack("<I", magic))
        if version is not None:
            fp.write(struct.pack("<I", version))
        fp.write(struct.pack("<7i", *hp))
        for i in range(hp[0]):
            t = (tokens or [b"a"] * hp[0])[i]
            fp.write(struct.pack("<I", len(t)) + t)
            if magic in (0x67676a74, 0x67676d66):
                fp.write(struct.pack("<f", 0.5))
        for n_dims, name, ftype, dims, payload in tensors:
            fp.write(struct.pack("<iiI", n_dims, len(name), ftype))
            fp.write(struct.pack(f"<{len(dims)}i", *dims))
            fp.write(name)
            if magic in (0x67676a74, 0x67676c61):
                fp.write(b"\0" * ((-fp.tell()) % 32))
            fp.write(payload)


@pytest.mark.parametrize("case,kind", [
    ("bad_magic", "InvalidMagic"), ("ggjt_v4", "InvalidFormatVersion"), ("ggmf_v2", "InvalidFormatVersion"), ("three_dims", "InvariantBroken"),
    ("q4_row_not_64", "InvariantBroken"), ("unknown_type", "UnsupportedElementType"), ("truncated_tensor", "Io"), ("truncated_header", "Io"),
    ("missing", "Io"),
])
def test_load_errors(tmp_path, case, kind):
    from llm_b200 import loader
    path = str(tmp_path / f"{case}.bin")
    f32 = lambda n: np.zeros(n, np.float32).tobytes()
    if case == "bad_magic":
        _raw_file(path, magic=0x12345678)
    elif case == "ggjt_v4":
        _raw_file(path, version=4)
    elif case == "ggmf_v2":
        _raw_file(path, magic=0x67676d66, version=2)
    elif case == "three_dims":
        _raw_file(path, tensors=[(3, b"x", 0, (2, 2, 2), f32(8))])
    elif case == "q4_row_not_64":
        _raw_file(path, tensors=[(2, b"w", 2, (32, 2), b"\0" * 36)])                 # loader.rs:249-254
    elif case == "unknown_type":
        _raw_file(path, tensors=[(1, b"k", 5, (256,), b"\0" * 144)])                 # 5 = the removed Q4_2: not an ElementType of the reference either
    elif case == "truncated_tensor":
        _raw_file(path, tensors=[(1, b"x", 0, (64,), f32(10))])
    elif case == "truncated_header":
        _raw_file(path)
        with open(path, "r+b") as fp:
            fp.truncate(20)
    with pytest.raises(loader.LoadError) as e:
        loader.GgmlFile(path)
    assert e.value.kind == kind


def test_legacy_containers_and_quantization_version_rule(orc, tmp_path):
    """unversioned 'ggml' (no scores, no alignment) and 'ggmf' v1 parse; GGJT v3 with file_type < 1000 is read as quantization version 2,
    GGJT v2 as 1 -> quantized tensors are refused (crates/llm-base/src/loader.rs:459-484)."""
    from llm_b200 import loader
    w = orc.quantize(B.Q4_0, np.ones((2, 64), np.float32)).tobytes()
    tensors = [(2, b"layers.0.attention.wq.weight", 2, (64, 2), w), (2, b"layers.0.feed_forward.w1.weight", 2, (64, 2), w), (1, b"norm.weight", 0, (64,), np.ones(64, np.float32).tobytes())]
    p = str(tmp_path / "legacy.bin")
    _raw_file(p, magic=0x67676d6c, version=None, tensors=tensors)
    f = loader.GgmlFile(p)
    assert f.container == ("ggml", 0) and [s for _, s in f.vocabulary()] == [0.0] * 4
    # not aligned in this container; the payload must still be found:
    assert np.array_equal(f.tensor_bytes(0), np.frombuffer(w, np.uint8)) and np.array_equal(f.tensor_bytes(2).view(np.float32), np.ones(64, np.float32))
    f.close()
    _raw_file(p, magic=0x67676d66, version=1, tensors=tensors)
    assert loader.GgmlFile(p).container == ("ggmf", 1)
    _raw_file(p, version=3, hp=(4, 64, 256, 2, 1, 32, 2), tensors=tensors)           # qnt version 0 in the file, GGJT v3 -> 2
    assert loader.GgmlFile(p).llama_hyperparameters()["quantization_version"] == 2
    _raw_file(p, version=2, hp=(4, 64, 256, 2, 1, 32, 2), tensors=tensors)           # GGJT v2 -> 1 -> refused
    with pytest.raises(loader.LoadError) as e:
        loader.GgmlFile(p).llama_hyperparameters()
    assert e.value.kind == "QuantizationVersion"
    _raw_file(p, version=3, hp=(4, 64, 256, 2, 1, 32, 1002), tensors=tensors)        # explicit version 1 -> refused
    with pytest.raises(loader.LoadError):
        loader.GgmlFile(p).llama_hyperparameters()


def test_gpt2_and_neox_headers(orc, tmp_path):
    """the other two architectures of the reference: 6 words + n_vocab repeated (gpt2 lib.rs:394-416), 8 words with a 0|1 bool (gptneox lib.rs:431-442)"""
    from llm_b200 import loader
    hp, tens = synth.make_gpt2(synth.GPT2_CONFIGS["gpt2-tiny"], B.Q4_0, orc.quantize, lm_head=True)
    p = str(tmp_path / "gpt2.ggjt")
    loader.write_model(p, "gpt2", dict(hp, file_type=2002), tens, synth.gpt2_tensor_shapes(hp, lm_head=True), hp["wtype"], vocabulary=_vocab(hp["n_vocab"]))
    f = loader.GgmlFile(p, "gpt2")
    h = f.hyperparameters()
    assert h == dict(n_vocab=hp["n_vocab"], n_ctx=hp["n_ctx"], n_embd=hp["n_embd"], n_head=hp["n_head"], n_layer=hp["n_layer"], file_type=2002, n_vocab_again=hp["n_vocab"])
    t = f.tensors()
    assert [x["name"] for x in t] == list(tens.keys()) and all(x["offset"] % 32 == 0 for x in t)
    wpe = next(i for i, x in enumerate(t) if x["name"] == "model/wpe")
    assert t[wpe]["type"] == 0 and np.array_equal(f.tensor_bytes(wpe).view(np.float32).reshape(hp["n_ctx"], hp["n_embd"]), tens["model/wpe"])
    f.close()
    with pytest.raises(loader.LoadError):                      # read as the wrong architecture: the vocabulary no longer lines up
        loader.GgmlFile(p, "gptneox")
    hn, tn = synth.make_neox(synth.NEOX_CONFIGS["neox-tiny"], B.Q5_1, orc.quantize)
    p2 = str(tmp_path / "neox.ggjt")
    loader.write_model(p2, "gptneox", dict(hn, file_type=2009), tn, synth.neox_tensor_shapes(hn), hn["wtype"])
    f = loader.GgmlFile(p2, "gptneox")
    h = f.hyperparameters()
    assert h["n_rot"] == hn["n_rot"] and h["use_parallel_residual"] == 1 and h["file_type"] == 2009 and h["n_ctx"] == hn["n_ctx"]
    assert len(f.tensors()) == len(tn) and len(f.vocabulary()) == hn["n_vocab"]
    f.close()
    # gpt2 header whose repeated n_vocab disagrees -> InvariantBroken; neox header whose bool is 2 -> Io (InvalidData)
    _raw_file(str(tmp_path / "bad1.bin"), hp=(4, 16, 64, 2, 1, 2002, 5))
    with pytest.raises(loader.LoadError) as e:
        loader.GgmlFile(str(tmp_path / "bad1.bin"), "gpt2")
    assert e.value.kind == "InvariantBroken"
    with open(str(tmp_path / "bad2.bin"), "wb") as fp:
        fp.write(struct.pack("<II8i", 0x67676a74, 3, 4, 16, 64, 2, 1, 8, 2, 2002))
    with pytest.raises(loader.LoadError) as e:
        loader.GgmlFile(str(tmp_path / "bad2.bin"), "gptneox")
    assert e.value.kind == "Io"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["q4_0", "q5_1", "q8_0"])
def test_model_loaded_from_file_is_bit_exact(orc, tmp_path, name):
    """llm::load::<Llama>(path): file -> mapping -> HBM; logits of prefill + decode equal the oracle's on the same tensors, bit for bit."""
    import llm_b200
    from llm_b200 import loader
    t = B.QUANT_TYPES[name]
    hp, tens, shapes = _model(orc, t, "tiny8")
    path = str(tmp_path / f"tiny8_{name}.ggjt")
    loader.write_llama(path, hp, tens, shapes, vocabulary=_vocab(hp["n_vocab"]))
    m = loader.load(path, llm_b200.ModelParameters(context_size=hp["n_ctx"]))
    assert m.hyperparameters["n_ff"] == hp["n_ff"] and m.hyperparameters["wtype"] == t
    s = m.start_session(llm_b200.InferenceSessionConfig(n_batch=64))
    mo = orc.llama(hp, tens)
    toks = synth.make_tokens(hp, 30)
    got, want = s.evaluate(toks[:24], all_logits=True), mo.eval(toks[:24])
    assert np.array_equal(np.asarray(got, np.float32).view(np.uint32).ravel(), np.asarray(want, np.float32).view(np.uint32).ravel())
    for i in range(24, 28):
        got, want = s.evaluate(toks[i:i + 1], all_logits=True), mo.eval(toks[i:i + 1])
        assert np.array_equal(np.asarray(got, np.float32).view(np.uint32).ravel(), np.asarray(want, np.float32).view(np.uint32).ravel()), i
    s.close(); m.close()


@pytest.mark.gpu
def test_load_file_rejects_incomplete_and_mismatched_models(orc, tmp_path):
    import llm_b200
    from llm_b200 import loader
    hp, tens, shapes = _model(orc, B.Q4_0, "tiny8")
    missing = {k: v for k, v in tens.items() if k != "layers.1.ffn_norm.weight"}
    p = str(tmp_path / "missing.ggjt")
    loader.write_llama(p, hp, missing, shapes)
    with pytest.raises(loader.LoadError) as e:
        loader.load(p, llm_b200.ModelParameters(context_size=hp["n_ctx"]))
    assert e.value.kind == "NotLoaded"
    extra = dict(tens); extra["layers.0.attention.bogus.weight"] = tens["layers.0.attention.wq.weight"]
    sh = dict(shapes); sh["layers.0.attention.bogus.weight"] = shapes["layers.0.attention.wq.weight"]
    loader.write_llama(p, hp, extra, sh)
    with pytest.raises(loader.LoadError) as e:
        loader.load(p, llm_b200.ModelParameters(context_size=hp["n_ctx"]))
    assert e.value.kind == "UnknownTensor"


def test_k_quant_tensors_parse_like_the_reference(tmp_path):
    """the reference's Type::try_from accepts the K-quant element types (crates/ggml/src/lib.rs:156-230): a file carrying them parses, with the block
    sizes of LC/k_quants.h (QK_K = 256); only LOADING such a tensor into this backend is refused"""
    from llm_b200 import loader
    path = str(tmp_path / "kq.bin")
    _raw_file(path, tensors=[(1, b"q4k", 12, (256,), b"\0" * 144), (1, b"q6k", 14, (512,), b"\0" * 420), (1, b"q2k", 10, (256,), b"\0" * 84)])
    f = loader.GgmlFile(path)
    infos = {t["name"]: t for t in f.tensors()}
    assert infos["q4k"]["nbytes"] == 144 and infos["q6k"]["nbytes"] == 420 and infos["q2k"]["nbytes"] == 84
