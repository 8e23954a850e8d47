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
