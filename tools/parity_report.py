#!/usr/bin/env python
"""GPU: measured parity margins of the tuned kernel variants against the golden rows of the
unmodified reference (tests/golden/boss_streaming_points.npz): multipoles inf-norm-relative and
elementwise, chi2 / lnL absolute.  Prints one JSON line per variant."""
import json
import os
import sys

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from victor_b200 import CCFFit  # noqa: E402

with open(os.path.join(ROOT, "config", "boss_config.yaml")) as fh:
    info = yaml.full_load(fh)
info["model"]["dir"] = info["data"]["dir"] = ROOT
fit = CCFFit(info["model"], info["data"], device=0)
g = np.load(os.path.join(ROOT, "tests", "golden", "boss_streaming_points.npz"))
eng, _ = fit._fit_engine({})
defaults = {"fast_math": 1, "newton": 3, "exp_degree": 5, "ilp": 4}
for opts in ({}, {"fast_math": 0}, {"exp_degree": 6}, {"newton": 2}, {"ilp": 1}):
    for k, v in {**defaults, **opts}.items():
        eng.set_option(k, v)
    lnl, chi2, th = fit.log_likelihood_batch(g["params"], return_theory=True)
    want = g["theory"]
    rel = max((np.abs(th[:, a:a + 30] - want[:, a:a + 30]).max(axis=1) / np.abs(want[:, a:a + 30]).max(axis=1)).max()
              for a in (0, 30))
    elem = np.abs(th - want) / (1e-9 * np.abs(want) + 1e-13)
    print(json.dumps({"variant": opts or "default", "multipole_infnorm_rel": float(rel),
                      "multipole_elementwise_over_tolerance": float(elem.max()),
                      "chi2_abs": float(np.abs(chi2 - g["chi2"]).max()), "lnl_abs": float(np.abs(lnl - g["lnl"]).max())}))
fit.close()
