#!/usr/bin/env python
"""GPU: 200,000 one- and two-row likelihood calls with random prior-box rows through the one-launch kernel (ticket
counters, mapped results, completion flags, graph replay), interleaved with batch calls; every result is compared with
the batch kernels' value for the same row."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import boss_blocks, synthetic_batch  # noqa: E402
from victor_b200 import CCFFit  # noqa: E402

fit = CCFFit(*boss_blocks(), device=0)
rows = synthetic_batch(65536, seed=77)
lnl_b, chi2_b = fit.log_likelihood_batch(rows)
t0 = time.time()
worst = 0.0
n_calls = 0
rng = np.random.default_rng(1)
i = 0
while n_calls < 200000:
    n = 1 + (n_calls % 7 == 0)
    i = int(rng.integers(0, len(rows) - 2))
    lnl, chi2 = fit.log_likelihood_batch(rows[i:i + n])
    worst = max(worst, float(np.max(np.abs(chi2 - chi2_b[i:i + n]))))
    n_calls += 1
    if n_calls % 50000 == 0:
        fit.log_likelihood_batch(rows[:5000])          # a batch call in between (other kernels, other graph state)
        print(n_calls, "calls", f"{time.time() - t0:.1f} s", "worst |chi2 - batch|", worst, flush=True)
assert worst < 1e-8, worst
print("ok:", n_calls, "calls,", f"{(time.time() - t0) / n_calls * 1e6:.1f} us per call incl. Python, worst chi2 difference to the batch path", worst)
fit.close()
