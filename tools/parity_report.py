#!/usr/bin/env python
"""GPU: measured parity margins of the tuned kernel variants
  (a) against the golden rows of the unmodified reference (tests/golden/boss_streaming_points.npz), and
  (b) against the C table walk over the first 16,384 rows of the bench batch:
multipoles inf-norm-relative (contract 1e-9) and the largest absolute error, chi2 / lnL absolute (contract 1e-6).
Prints one JSON line per variant."""
import json
import os
import sys

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synthetic_batch  # noqa: E402
from oracle.table_walk import TableWalk  # noqa: E402
from victor_b200 import CCFFit  # noqa: E402
from victor_b200.model import params_to_rows  # noqa: E402

with open(os.path.join(ROOT, "config", "boss_config.yaml")) as fh:
    info = yaml.full_load(fh)
info["model"]["dir"] = info["data"]["dir"] = ROOT
fit = CCFFit(info["model"], info["data"], device=0)
g = np.load(os.path.join(ROOT, "tests", "golden", "boss_streaming_points.npz"))
rows = params_to_rows(synthetic_batch(65536)[:16384])


def margins(th, want, ns=30):
    th, want = th.reshape(len(th), -1, ns), want.reshape(len(want), -1, ns)
    rel = (np.abs(th - want).max(axis=2) / np.abs(want).max(axis=2)).max()
    return float(rel), float(np.abs(th - want).max())


for kw in ({}, {"rsd_model": "dispersion"}, {"assume_isotropic": False}):
    eng, _ = fit._fit_engine(kw)
    wth, wc2, wll = TableWalk(fit, options=kw).likelihood(rows, want_theory=True)
    variants = [{"fast_math": 0}, {"newton": 3, "exp_degree": 5}, {"newton": 2, "exp_degree": 5}]
    if not kw:
        variants += [{"newton": 2, "exp_degree": 3}]
    else:
        variants = [{"fast_math": 0}, {"tuned": 0}, {"tuned": 1}]
    for opts in variants:
        for k, v in {"fast_math": 1, "ilp": 0, "tuned": 1, "newton": 0, "exp_degree": 0, **opts}.items():
            eng.set_option(k, v)
        out = {"model": kw or "streaming", "variant": opts}
        if not kw:
            lnl, chi2, th = fit.log_likelihood_batch(g["params"], return_theory=True)
            rel, ab = margins(th, g["theory"])
            out["golden_80_rows"] = {"multipole_infnorm_rel": rel, "multipole_abs": ab,
                                     "chi2_abs": float(np.abs(chi2 - g["chi2"]).max()),
                                     "lnl_abs": float(np.abs(lnl - g["lnl"]).max())}
        lnl, chi2, th = fit.log_likelihood_batch(rows, return_theory=True, **kw)
        rel, ab = margins(th, wth)
        out["table_walk_16384_rows"] = {"multipole_infnorm_rel": rel, "multipole_abs": ab,
                                        "chi2_abs": float(np.nanmax(np.abs(chi2 - wc2))),
                                        "lnl_abs": float(np.nanmax(np.abs(lnl - wll)))}
        print(json.dumps(out), flush=True)
fit.close()
