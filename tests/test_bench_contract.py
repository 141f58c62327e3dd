"""CPU: the JSON line of ``bench.py --impl reference`` (the one arm that runs without a GPU) carries the keys
the driver reads, on this repository's metric / unit / config."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_line(env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, timeout=600,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def _check_reference_line(d, kind):
    with open(os.path.join(ROOT, "BASELINE.json")) as fh:
        base = json.load(fh)
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["metric"].split(",")[0] == base["metric"].split(",")[0]        # "likelihood evals/sec (multipoles+chi2)"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "BOSS DR12 CMASS" in d["config"]["workload"] and d["config"]["rows_per_step"] >= 4
    assert f"{d['config']['rows_per_step']}-row sample" in d["config"]["workload"]     # why same_config is false
    cb = d["cpu_baseline"]
    assert cb["kind"] == kind and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_line():
    """The arm drives the UNMODIFIED reference (baseline/_ref or the dev container's checkout) when one is
    importable, else the oracle port, and says which."""
    from oracle import refshim
    kind = "reference" if refshim.find_reference() else "port"
    d = _reference_line()
    _check_reference_line(d, kind)
    if kind == "reference":
        assert "UNMODIFIED reference" in d["cpu_baseline"]["sample"]


def test_reference_arm_falls_back_to_the_port():
    """Without an importable reference (VB200_NO_REFERENCE hides it, as on a box that got no baseline/_ref) the
    same command times the oracle port and labels it so."""
    d = _reference_line({"VB200_NO_REFERENCE": "1"})
    _check_reference_line(d, "port")


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_line():
    """The default arm on one GPU (short run, CPU leg skipped): one JSON line on stdout with the contract's keys."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "3", "--no-cpu",
                          "--sweep", "131072", "--sustain", "0"],
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
    assert d["value"] > 1e6 and d["gpu_launches"] == 8                    # K1 + K2 (bracket count, scatter, chi2) per step
    x = d["extra"]
    for name in ("dispersion", "dense_sweep", "mcmc", "strong_64k"):
        assert x[name]["value"] > 0, name
    assert x["dispersion"]["roofline"]["flop_per_point"] == 117 and 0.1 < x["dispersion"]["roofline"]["frac"] < 1.0
    assert x["dense_sweep"]["scaling"] == "strong" and x["dense_sweep"]["rows_total"] == 131072
    assert x["mcmc"]["unit"] == "calls/s" and x["mcmc"]["latency_us"]["median"] > 0
    assert x["strong_64k"]["rows_total"] == 65536 and x["strong_64k"]["all_rows_finite"]
    assert "sustained" not in x
    e = d["e2e"]
    assert 0 < e["value"] <= 1.05 * d["value"] and e["h2d_bytes_per_step"] == 65536 * 80 and e["d2h_bytes_per_step"] == 65536 * 16
    r = d["roofline"]
    assert r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0.2 < r["frac"] < 1.0
    assert abs(r["peak"] - 37.2) < 0.1 and "nominal" in r["peak_source"]        # quoted against the nominal FP64 peak
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["clocks"]
    assert c["sm_max_mhz"] >= c["sm_mhz"] > 0 and isinstance(c["reasons"], list)
