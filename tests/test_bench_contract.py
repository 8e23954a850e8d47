"""bench.py contract, CPU side: the reference arm (`--impl reference`) prints exactly one JSON line on stdout with the keys the driver reads,
times the reference's own ggml CPU path (oracle/_ref when built, else the plain-C port) and needs no GPU.  The contract test runs the
2-layer debug size (--layers 2); the driver's run uses the full 32-layer model."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--layers", "2"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "tokens/s" and j["higher_is_better"] is True and j["scaling"] == "weak"
    assert j["metric"].startswith("LLaMA-7B Q4_0 tokens/sec") and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["n_gpus"] == 1
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == j["value"] and "7B" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]


def test_reference_arm_prefill_metric():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--metric", "prefill", "--layers", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    j = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
    assert j["impl"] == "reference" and "prefill@512" in j["metric"] and j["value"] > 0 and j["cpu_baseline"]["value"] == j["value"]


def test_clock_sampler_holds_the_load_until_clocks_were_seen():
    """ClockSampler.hold: bursts run until nvidia-smi has reported (here: a fake pump), every rank runs the count the `agree` hook returns, and a host
    without nvidia-smi (proc None) runs none"""
    sys.path.insert(0, ROOT)
    import bench
    row = ["0", "1965", "1965", "300", "Not Active", "Not Active", "Not Active", "Active"]
    c = bench.ClockSampler(0); c.proc = object(); n = [0]

    def burst():
        n[0] += 1
        if n[0] == 3:
            c.rows += [row, row]
    c.hold(burst)
    assert n[0] == 3
    s = c.summary()
    assert s["samples"] == 2 and s["sm_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"]
    # another rank still needs samples: this rank keeps bursting although it has its own
    calls = []
    c.hold(lambda: calls.append(1), agree=lambda need: 1.0 if len(calls) < 2 else need)
    assert len(calls) == 2
    c2 = bench.ClockSampler(0); c2.proc = None
    c2.hold(burst)
    assert n[0] == 3
    # nothing ever reported: bounded by max_s
    c3 = bench.ClockSampler(0); c3.proc = object(); k = [0]
    c3.hold(lambda: k.__setitem__(0, k[0] + 1), max_s=0.05)
    assert k[0] >= 1 and c3.summary()["samples"] == 0
