"""CPU: the JSON line of ``bench.py --impl reference`` (the one arm that runs without a GPU) carries the keys
the driver reads, on this repository's metric / unit / config."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        base = json.load(fh)
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["metric"].split(",")[0] == base["metric"].split(",")[0]        # "likelihood evals/sec (multipoles+chi2)"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "BOSS DR12 CMASS" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_line():
    """The default arm on one GPU (short run, CPU leg skipped): one JSON line on stdout with the contract's keys."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu"],
                         capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1 and lines[0].startswith("{"), out.stdout[:500]
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in d, key
    assert d["unit"] == "evals/s" and d["n_gpus"] == 1 and d["steps"] == 2 and d["scaling"] == "weak"
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "BOSS DR12 CMASS" in d["config"]["workload"] and "l2" in d["config"]
    assert d["value"] > 1e6 and d["gpu_launches"] == 4                    # K1 + K2 per step
    e = d["e2e"]
    assert 0 < e["value"] <= 1.05 * d["value"] and e["h2d_bytes_per_step"] == 65536 * 80 and e["d2h_bytes_per_step"] == 65536 * 16
    r = d["roofline"]
    assert r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.2 < r["frac"] < 1.0
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["clocks"]
    assert c["sm_max_mhz"] >= c["sm_mhz"] > 0 and isinstance(c["reasons"], list)
