"""Tensor-parallel host logic on CPU (no GPU): the row-split rules of llm_b200/tp.py, alone and across two gloo ranks launched by torchrun -- the same
shard_tensors / all_gather_object path the GPU ranks use to cut their shards and to exchange the 64-byte CUDA IPC handles of their exchange slabs."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT
from oracle import bindings as B
from oracle import synth


def test_row_split_rules_reassemble_and_match_the_oracle(orc):
    """each rank's rows of wq/wk/wv/wo/w1/w2/w3/output are disjoint, cover the tensor, and a mat-vec on a shard equals the matching rows of the
    full mat-vec bit for bit (the N-split is exact by construction: one complete vec_dot per dst element)"""
    from llm_b200 import tp
    hp, tens = synth.make_llama(dict(synth.CONFIGS["gqa8"], n_layer=1, n_vocab=384), B.Q5_1, orc.quantize)
    for world in (2,):
        shards = [tp.shard_tensors(hp, tens, r, world) for r in range(world)]
        for name, full in tens.items():
            back = tp.unshard_rows(hp, name, [s[name] for s in shards])
            assert back.shape == full.shape and np.array_equal(back, full), name
        x = np.random.default_rng(3).standard_normal((3, hp["n_embd"])).astype(np.float32)
        name = "layers.0.attention.wo.weight"
        whole = orc.mul_mat(B.Q5_1, tens[name], x)
        parts = np.concatenate([orc.mul_mat(B.Q5_1, s[name], x) for s in shards], axis=1)
        assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32))
    with pytest.raises(ValueError):
        tp.shard_tensors(hp, tens, 0, 3)                    # 8 heads / 2 KV heads do not split three ways


def test_row_split_at_4_and_8_ranks_and_published_geometries(orc):
    """the same reassembly + exactness at 4 and 8 ranks (the multi-head test geometry of tests/test_gpu_tp.py), and the divisibility rule on the published
    shapes: LLaMA-7B and 13B split 2 / 4 / 8 ways (whole heads, 32-row pieces of w1|w3 / wo / w2 / lm_head), not 3 or 16 ways"""
    from llm_b200 import tp
    cfg = dict(n_vocab=1024, n_embd=512, n_head=8, n_head_kv=8, n_layer=1, n_ff=1024, n_rot=64, n_ctx=64)
    hp, tens = synth.make_llama(cfg, B.Q4_0, orc.quantize)
    x = np.random.default_rng(4).standard_normal((2, hp["n_embd"])).astype(np.float32)
    for world in (4, 8):
        shards = [tp.shard_tensors(hp, tens, r, world) for r in range(world)]
        for name, full in tens.items():
            assert np.array_equal(tp.unshard_rows(hp, name, [s[name] for s in shards]), full), (world, name)
        for name in ("layers.0.feed_forward.w1.weight", "output.weight"):
            whole = orc.mul_mat(B.Q4_0, tens[name], x)
            parts = np.concatenate([orc.mul_mat(B.Q4_0, s[name], x) for s in shards], axis=1)
            assert np.array_equal(whole.view(np.uint32), parts.view(np.uint32)), (world, name)
    for geom in ("7b", "13b"):
        g = synth.CONFIGS[geom]
        for world in (2, 4, 8):
            tp.check_divisible(g, world)
        for world in (3, 16):
            with pytest.raises(ValueError):
                tp.check_divisible(g, world)


WORKER = textwrap.dedent('''
    import json, os, sys
    sys.path.insert(0, %r)
    import numpy as np
    import torch.distributed as dist
    from oracle import bindings as B
    from oracle import synth
    from llm_b200 import tp
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    orc = B.Oracle()
    hp, tens = synth.make_llama(dict(synth.CONFIGS["tiny"], n_layer=1), B.Q4_0, orc.quantize)
    mine = tp.shard_tensors(hp, tens, rank, world)
    # what TpSession does with the IPC handles: every rank contributes 64 bytes, every rank gets the table in rank order
    table = [None] * world
    dist.all_gather_object(table, bytes([rank]) * 64)
    assert [t[0] for t in table] == list(range(world)) and all(len(t) == 64 for t in table)
    # the decode exchange in host form: each rank computes ITS rows of wo x, the slices concatenated in rank order are the full result
    x = np.random.default_rng(5).standard_normal((1, hp["n_embd"])).astype(np.float32)
    part = orc.mul_mat(B.Q4_0, mine["layers.0.attention.wo.weight"], x)
    parts = [None] * world
    dist.all_gather_object(parts, part)
    full = orc.mul_mat(B.Q4_0, tens["layers.0.attention.wo.weight"], x)
    ok = bool(np.array_equal(np.concatenate(parts, axis=1).view(np.uint32), full.view(np.uint32)))
    rows = {k: list(tp.shard_rows(k, hp, rank, world) or ()) for k in ("layers.0.attention.wq.weight", "layers.0.feed_forward.w2.weight", "output.weight", "norm.weight")}
    open(os.path.join(os.environ["RESULT_DIR"], "rank%%d.json" %% rank), "w").write(json.dumps({"rank": rank, "ok": ok, "rows": rows}))
    dist.barrier()
    dist.destroy_process_group()
''') % ROOT


def test_two_rank_gloo_shards_and_handle_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1", RESULT_DIR=str(tmp_path))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29519", str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    res = sorted((json.loads((tmp_path / f"rank{i}.json").read_text()) for i in range(2)), key=lambda r: r["rank"])
    assert all(r["ok"] for r in res)
    e, v = synth.CONFIGS["tiny"]["n_embd"], synth.CONFIGS["tiny"]["n_vocab"]
    assert res[0]["rows"]["layers.0.attention.wq.weight"] == [0, e // 2] and res[1]["rows"]["layers.0.attention.wq.weight"] == [e // 2, e]
    assert res[1]["rows"]["output.weight"] == [v // 2, v] and res[0]["rows"]["norm.weight"] == []
