"""Tensor-parallel host layer (one process per GPU, launched by torchrun): shards a LLaMA model by OUTPUT ROWS across the ranks and wires the
ranks' exchange slabs together (include/llm_b200.h: b200_llama_new_tp, b200_session_tp_handle / _connect; kernels: llm_b200/csrc/tp.cuh).

Row split, per tensor (G ranks, rank r) -- every dst element stays ONE complete ggml_vec_dot over the full K (LC/ggml.c:10570-10572), so the result is
bit-identical to the single-GPU / CPU result:
    attention.wq          rows [r e/G, (r+1) e/G)          = the rank's heads
    attention.wk / wv     rows [r gqa/G, (r+1) gqa/G)      = the rank's KV heads (its slice of the KV cache)
    attention.wo          rows [r e/G, ...)                 output channels; input = all heads' attention rows (gathered by the attention epilogues)
    feed_forward.w1 / w3  rows [r f/G, ...)
    feed_forward.w2       rows [r e/G, ...)                 input = all ranks' silu(w1 x) * (w3 x) slices
    output                rows [r V/G, ...)
    tok_embeddings, norms whole on every rank
torch.distributed carries only set-up traffic (the 64-byte CUDA IPC handles) and the barriers around timed regions; the per-token exchange is
done by the decode kernels themselves (peer stores over NVLink as tagged 8-byte units, polled locally by the consumers: tp.cuh).  Contrast: LC/ggml-cuda.cu:3355-3583."""
import ctypes as C
import os
from typing import Dict

import numpy as np

from . import _lib
from .session import InferenceSession, InferenceSessionConfig, Llama, ModelParameters, _HP_KEYS, _check


def shard_rows(name: str, hp: Dict[str, int], rank: int, world: int):
    """(row0, row1) of the full tensor `name` owned by `rank`, or None when the tensor is replicated."""
    e, f, v = hp["n_embd"], hp["n_ff"], hp["n_vocab"]
    gqa = e // (hp["n_head"] // hp["n_head_kv"])
    if name in ("tok_embeddings.weight", "norm.weight") or name.endswith("norm.weight"):
        return None
    if name == "output.weight":
        n = v
    else:
        sub = name.split(".", 2)[2]
        n = {"attention.wq.weight": e, "attention.wk.weight": gqa, "attention.wv.weight": gqa, "attention.wo.weight": e,
             "feed_forward.w1.weight": f, "feed_forward.w3.weight": f, "feed_forward.w2.weight": e}[sub]
    assert n % world == 0, (name, n, world)
    return rank * (n // world), (rank + 1) * (n // world)


def check_divisible(hp: Dict[str, int], world: int):
    e, f, v = hp["n_embd"], hp["n_ff"], hp["n_vocab"]
    ok = (hp["n_head"] % world == 0 and hp["n_head_kv"] % world == 0 and f % world == 0 and (f // world) % 32 == 0 and (e // world) % 32 == 0
          and v % world == 0 and (v // world) % 32 == 0)
    if not ok:
        raise ValueError(f"tensor-parallel degree {world} does not divide the model: heads {hp['n_head']}/{hp['n_head_kv']}, n_ff {f}, n_embd {e}, n_vocab {v} "
                         "(whole heads per rank; 32-row pieces of w1|w3, wo/w2 and the lm_head)")


def shard_tensors(hp: Dict[str, int], tensors: Dict[str, np.ndarray], rank: int, world: int) -> Dict[str, np.ndarray]:
    """Full tensors (uint8 block rows [N, row_bytes] / f32 vectors) -> this rank's shards.  Pure host logic (tests/test_tp_gloo.py)."""
    check_divisible(hp, world)
    out = {}
    for name, arr in tensors.items():
        rows = shard_rows(name, hp, rank, world)
        out[name] = arr if rows is None else np.ascontiguousarray(arr[rows[0]:rows[1]])
    return out


def unshard_rows(hp: Dict[str, int], name: str, shards):
    """inverse of shard_tensors for one tensor: concatenate the ranks' shards (tests)"""
    return shards[0] if shard_rows(name, hp, 0, len(shards)) is None else np.concatenate(shards, axis=0)


class TpLlama(Llama):
    """Llama whose weights are the `rank`-th row shard (KnownModel over b200_llama_new_tp)."""

    def __init__(self, hyperparameters, params: ModelParameters = None, tensors=None, rank=0, world=1, device=None, presharded=False):
        params = params or ModelParameters()
        check_divisible(hyperparameters, world)
        self.L = _lib.lib()
        _check(self.L.b200_init(rank if device is None else device), "b200_init")
        self.hyperparameters = dict(hyperparameters)
        self.params, self.rank, self.world = params, rank, world
        hp = _lib.LlamaHparams(**{k: int(hyperparameters[k]) for k in _HP_KEYS}, wtype=int(hyperparameters["wtype"]),
                               context_size=params.context_size, rope_freq_base=params.rope_freq_base, rope_freq_scale=params.rope_freq_scale)
        self._m = self.L.b200_llama_new_tp(C.byref(hp), rank, world)
        if not self._m:
            raise ValueError(f"b200_llama_new_tp rejected {hyperparameters} at degree {world}")
        if tensors is not None:
            local = tensors if presharded else shard_tensors(hyperparameters, tensors, rank, world)
            for name, arr in local.items():
                self.load_tensor(name, arr)

    def start_session(self, config: InferenceSessionConfig = None, dist=None) -> "TpSession":
        return TpSession(self, config or InferenceSessionConfig(), dist)


class TpSession(InferenceSession):
    """InferenceSession of a tensor-parallel shard.  Every rank calls evaluate() with the same tokens; logits are complete on every rank."""

    def __init__(self, model: TpLlama, config: InferenceSessionConfig, dist=None):
        super().__init__(model, config)
        self.rank, self.world = model.rank, model.world
        handle = (C.c_ubyte * 64)()
        _check(self.L.b200_session_tp_handle(self._s, handle), "tp_handle")
        mine = bytes(handle)
        if dist is None:
            import torch.distributed as dist
        table = [None] * self.world
        dist.all_gather_object(table, mine)                 # set-up traffic only
        blob = b"".join(table)
        _check(self.L.b200_session_tp_connect(self._s, blob), "tp_connect")
        dist.barrier()                                      # every rank has mapped every slab before the first token

    def kv(self, which: int) -> np.ndarray:
        hp = self.model.hyperparameters
        gqa = hp["n_embd"] // (hp["n_head"] // hp["n_head_kv"]) // self.world
        out = np.empty(hp["n_layer"] * self.model.params.context_size * gqa, np.uint16)
        _check(self.L.b200_session_read_kv(self._s, which, out.ctypes.data_as(C.c_void_p), out.nbytes), "read_kv")
        return out

    @property
    def timeouts(self) -> int:
        return self.L.b200_session_tp_timeouts(self._s)


def init_distributed():
    """torchrun environment -> (rank, local_rank, world, dist); NCCL when a GPU is present (barriers / max-over-ranks), gloo otherwise"""
    rank, local_rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group("gloo")
    return rank, local_rank, world, dist
