#!/usr/bin/env python
"""End-to-end cost of bulk host-bound outputs (theory vectors): one launch + one copy against row chunks whose
copies overlap the computation of later chunks."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import DENSE, boss_blocks, synthetic_batch  # noqa: E402
from victor_b200 import CCFFit  # noqa: E402
from victor_b200.model import params_to_rows  # noqa: E402


def timeit(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts))


def main():
    model, data = boss_blocks()
    fit = CCFFit(model, data, device=0)
    rows = params_to_rows(synthetic_batch(65536))
    eng, _ = fit._fit_engine({})
    print("BOSS 65536 rows, lnL + chi2 only              ms:", timeit(lambda: fit.log_likelihood_batch(rows)))
    for chunks in (1, 0, 4, 8, 16):
        eng.set_option("chunks", chunks)
        print(f"BOSS 65536 rows, with theory vectors, chunks={chunks:2d} ms:",
              timeit(lambda: fit.log_likelihood_batch(rows, return_theory=True)))
    eng.set_option("chunks", 0)
    kw = dict(velocity_nodes=DENSE["nx"], mu_nodes=DENSE["nmu"])
    deng = fit._engine(fit._merged_options(kw))
    r2 = rows[:32768]
    for chunks in (1, 0, 8):
        deng.set_option("chunks", chunks)
        print(f"dense 32768 rows, l = 0, 2, 4 vectors, chunks={chunks:2d}  ms:",
              timeit(lambda: fit.theory_multipole_vector_batch(np.asarray(fit.s, float), r2, DENSE["poles"], **kw), reps=3))
    fit.close()


if __name__ == "__main__":
    main()
