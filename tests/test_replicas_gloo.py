"""N>1 host logic on CPU: two processes over gloo (torchrun), the same code path bench.py --gpus N uses for its barrier and its
max-over-ranks timing (the GPU work itself needs no collective: independent replicas)."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent('''
    import json, os, sys
    sys.path.insert(0, %r)
    from llm_b200.replicas import Replicas
    r = Replicas(backend="gloo")
    assert r.world == 2 and r.rank in (0, 1)
    r.barrier()
    ms = [10.0 + 5.0 * r.rank, 3.0 - r.rank]            # rank 1 is slower on the first timer, rank 0 on the second
    mx = r.max_over_ranks(ms)
    r.barrier()
    out = {"rank": r.rank, "max": mx, "rate": r.whole_job_rate(64, mx[0])}
    open(os.path.join(os.environ["RESULT_DIR"], "rank" + str(r.rank) + ".json"), "w").write(json.dumps(out))
    r.close()
''') % ROOT


def test_two_rank_gloo_barrier_and_max(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, OMP_NUM_THREADS="1", RESULT_DIR=str(tmp_path))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    res = [json.loads((tmp_path / f"rank{i}.json").read_text()) for i in range(2)]
    assert sorted(r["rank"] for r in res) == [0, 1]
    for r in res:
        assert r["max"] == [15.0, 3.0]                   # element-wise max over ranks, identical on every rank
        assert abs(r["rate"] - 2 * 64 / 15e-3) < 1e-6    # whole-job rate = all ranks' units / slowest rank's time


def test_single_process_is_a_noop():
    from llm_b200.replicas import Replicas
    r = Replicas()
    assert r.world == 1
    r.barrier()
    assert r.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]
    assert r.whole_job_rate(10, 5.0) == 2000.0
