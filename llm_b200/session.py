"""Host-side mirror of the reference's model/session interface for LLaMA over the native C ABI (include/llm_b200.h).

Names follow the reference so that the parity tests read like its own integration tests (binaries/llm-test):
  Llama(KnownModel)                crates/models/llama/src/lib.rs:17-140   new / start_session / evaluate
  ModelParameters                  crates/llm-base/src/model/mod.rs:197-229
  InferenceSessionConfig           crates/llm-base/src/inference_session.rs:799-841
  InferenceSession                 crates/llm-base/src/inference_session.rs:43-512  feed_prompt / n_past / last_logits
  OutputRequest                    crates/llm-base/src/lib.rs (all_logits), model/common.rs:22-39
  InferenceError::ContextFull      crates/llm-base/src/inference_session.rs:311-313
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np

from . import _lib


class ContextFull(Exception):
    """InferenceError::ContextFull."""


@dataclass
class ModelParameters:
    context_size: int = 2048          # model/mod.rs:214
    use_gpu: bool = True              # this backend IS the GPU path; False is rejected (no CPU fallback)
    rope_freq_base: float = 10000.0   # RoPEOverrides
    rope_freq_scale: float = 1.0


@dataclass
class InferenceSessionConfig:
    n_batch: int = 8                  # inference_session.rs:837
    flags: int = 0


@dataclass
class OutputRequest:
    all_logits: Optional[np.ndarray] = None     # set to an empty array to request all rows (OutputRequest::all_logits)
    want_all_logits: bool = False


_HP_KEYS = ("n_vocab", "n_embd", "n_head", "n_head_kv", "n_layer", "n_rot", "n_ff")


def llama_tensor_shapes(hp: Dict[str, int]) -> Dict[str, tuple]:
    """tensor name -> (rows N, cols K) for 2-D weights, (n,) for norm gains: the names the reference's loader asks for
    (crates/models/llama/src/lib.rs:52-91)"""
    e, f, v = hp["n_embd"], hp["n_ff"], hp["n_vocab"]
    gqa = e // (hp["n_head"] // hp["n_head_kv"])
    shapes = {"tok_embeddings.weight": (v, e), "norm.weight": (e,), "output.weight": (v, e)}
    for i in range(hp["n_layer"]):
        p = f"layers.{i}."
        shapes.update({p + "attention_norm.weight": (e,), p + "attention.wq.weight": (e, e), p + "attention.wk.weight": (gqa, e),
                       p + "attention.wv.weight": (gqa, e), p + "attention.wo.weight": (e, e), p + "ffn_norm.weight": (e,),
                       p + "feed_forward.w1.weight": (f, e), p + "feed_forward.w2.weight": (e, f), p + "feed_forward.w3.weight": (f, e)})
    return shapes


def _check(rc, what):
    if rc == -1:
        raise ContextFull(what)
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc}")


class Llama:
    """KnownModel for LLaMA.  `hyperparameters` is a dict with n_vocab, n_embd, n_head, n_head_kv, n_layer, n_rot, n_ff, wtype."""

    def __init__(self, hyperparameters: Dict[str, int], params: ModelParameters = None, tensors: Dict[str, np.ndarray] = None,
                 device: int = 0):
        params = params or ModelParameters()
        if not params.use_gpu:
            raise ValueError("llm_b200 has no CPU path: ModelParameters.use_gpu must be True")
        self.L = _lib.lib()
        _check(self.L.b200_init(device), "b200_init")
        self.hyperparameters = dict(hyperparameters)
        self.params = params
        hp = _lib.LlamaHparams(**{k: int(hyperparameters[k]) for k in _HP_KEYS}, wtype=int(hyperparameters["wtype"]),
                               context_size=params.context_size, rope_freq_base=params.rope_freq_base,
                               rope_freq_scale=params.rope_freq_scale)
        self._m = self.L.b200_llama_new(C.byref(hp))
        if not self._m:
            raise ValueError(f"b200_llama_new rejected hyperparameters {hyperparameters}")
        if tensors is not None:
            for name, arr in tensors.items():
                self.load_tensor(name, arr)

    # TensorLoader::load + transfer_to(Backend::Gpu)
    def load_tensor(self, name: str, arr: np.ndarray):
        arr = np.ascontiguousarray(arr)
        typ = 0 if arr.dtype == np.float32 else int(self.hyperparameters["wtype"])
        _check(self.L.b200_model_load_tensor(self._m, name.encode(), typ, arr.ctypes.data_as(C.c_void_p), arr.nbytes),
               f"load_tensor({name})")

    def synthesize(self, seed: int = 0x5EED0000):
        _check(self.L.b200_model_synthesize(self._m, seed), "synthesize")

    def read_tensor(self, name: str) -> np.ndarray:
        nb = self.L.b200_model_tensor_nbytes(self._m, name.encode())
        if nb == 0:
            raise KeyError(name)
        is_f32 = name.endswith("norm.weight")
        out = np.empty(nb // 4, np.float32) if is_f32 else np.empty(nb, np.uint8)
        _check(self.L.b200_model_read_tensor(self._m, name.encode(), out.ctypes.data_as(C.c_void_p), nb), f"read_tensor({name})")
        return out

    @property
    def weight_bytes(self) -> int:
        return self.L.b200_model_weight_bytes(self._m)

    def start_session(self, config: InferenceSessionConfig = None) -> "InferenceSession":
        return InferenceSession(self, config or InferenceSessionConfig())

    def context_size(self) -> int:
        return self.params.context_size

    def close(self):
        if getattr(self, "_m", None):
            self.L.b200_model_free(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class InferenceSession:
    def __init__(self, model: Llama, config: InferenceSessionConfig):
        self.model, self.config, self.L = model, config, model.L
        cfg = _lib.SessionConfig(n_batch=config.n_batch, flags=config.flags)
        self._s = self.L.b200_model_start_session(model._m, C.byref(cfg))
        if not self._s:
            raise RuntimeError("b200_model_start_session failed (all tensors loaded?)")
        self.n_vocab = int(model.hyperparameters["n_vocab"])
        self.last_logits = np.zeros(self.n_vocab, np.float32)

    @property
    def n_past(self) -> int:
        return self.L.b200_session_n_past(self._s)

    def top_k(self, k: int):
        """Sampler hand-off: (token ids, logits) of the k largest logits of the last evaluated row, selected on the device."""
        ids, vals = np.empty(k, np.int32), np.empty(k, np.float32)
        _check(self.L.b200_session_top_k(self._s, k, ids.ctypes.data_as(C.c_void_p), vals.ctypes.data_as(C.c_void_p)), "top_k")
        return ids, vals

    def evaluate(self, tokens, all_logits: bool = False) -> np.ndarray:
        """Model::evaluate: one forward pass over `tokens` (<= n_batch) appended at n_past; returns logits
        ([n, n_vocab] when all_logits else the last row, which is also kept in self.last_logits)."""
        tokens = np.ascontiguousarray(tokens, np.int32)
        n = tokens.size
        out = np.empty((n if all_logits else 1, self.n_vocab), np.float32)
        _check(self.L.b200_session_evaluate(self._s, tokens.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p),
                                            1 if all_logits else 0), "evaluate")
        self.last_logits = out[-1].copy()
        return out if all_logits else out[0]

    def feed_prompt(self, tokens) -> np.ndarray:
        tokens = np.ascontiguousarray(tokens, np.int32)
        if tokens.size == 0:                      # the reference leaves last_logits untouched for an empty prompt (inference_session.rs:299-350)
            return self.last_logits
        out = np.empty(self.n_vocab, np.float32)
        _check(self.L.b200_session_feed_prompt(self._s, tokens.ctypes.data_as(C.c_void_p), tokens.size,
                                               out.ctypes.data_as(C.c_void_p)), "feed_prompt")
        self.last_logits = out
        return out

    def rewind(self, n_past: int):
        _check(self.L.b200_session_set_n_past(self._s, n_past), "rewind")

    def kv(self, which: int) -> np.ndarray:
        hp = self.model.hyperparameters
        gqa = hp["n_embd"] // (hp["n_head"] // hp["n_head_kv"])
        n = hp["n_layer"] * self.model.params.context_size * gqa
        out = np.empty(n, np.uint16)
        _check(self.L.b200_session_read_kv(self._s, which, out.ctypes.data_as(C.c_void_p), out.nbytes), "read_kv")
        return out

    def sync(self):
        self.L.b200_session_sync(self._s)

    @property
    def last_launches(self) -> int:
        return self.L.b200_session_last_launches(self._s)

    def close(self):
        if getattr(self, "_s", None):
            self.L.b200_session_free(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
