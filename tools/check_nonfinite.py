#!/usr/bin/env python
"""GPU: rows of the 65,536-row bench batch whose dispersion-model likelihood is not finite, per kernel variant,
and what the C table walk (CPU) gives for exactly those rows."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import boss_blocks, synthetic_batch  # noqa: E402
from oracle.table_walk import TableWalk  # noqa: E402
from victor_b200 import CCFFit  # noqa: E402
from victor_b200.model import params_to_rows  # noqa: E402

fit = CCFFit(*boss_blocks(), device=0)
rows = params_to_rows(synthetic_batch(65536))
kw = {"rsd_model": "dispersion"}
eng, _ = fit._fit_engine(kw)
bad_all = set()
res = {}
for name, opts in (("tuned", {"tuned": 1, "fast_math": 1}), ("general", {"tuned": 0, "fast_math": 1}),
                   ("libm", {"tuned": 0, "fast_math": 0})):
    for k, v in opts.items():
        eng.set_option(k, v)
    lnl, chi2, th = fit.log_likelihood_batch(rows, return_theory=True, **kw)
    bad = np.flatnonzero(~np.isfinite(lnl))
    res[name] = (lnl, chi2, th)
    print(name, "non-finite rows:", len(bad), bad[:12].tolist())
    bad_all.update(bad.tolist())
eng.set_option("tuned", 1)
eng.set_option("fast_math", 1)
bad_all = sorted(bad_all)
if bad_all:
    sub = rows[bad_all]
    wth, wc2, wll = TableWalk(fit, options=kw).likelihood(sub, want_theory=True)
    for i, r in enumerate(bad_all):
        print("row", r, "params", rows[r, :5].tolist(), "C walk lnl", wll[i], "chi2", wc2[i], "theory nan:", int(np.isnan(wth[i]).sum()),
              {k: (float(v[0][r]), int(np.isnan(v[2][r]).sum())) for k, v in res.items()})
fit.close()
