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
        _raw_file(path, tensors=[(1, b"k", 12, (256,), b"\0" * 144)])                # a K-quant id: not an element type of this backend
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
