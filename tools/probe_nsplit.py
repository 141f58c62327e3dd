#!/usr/bin/env python
"""Throughput against batch size, automatic block split against forced splits (device-resident buffers)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import boss_blocks, synthetic_batch, P  # noqa: E402
from victor_b200 import CCFFit  # noqa: E402
from victor_b200.model import params_to_rows  # noqa: E402


def main():
    model, data = boss_blocks()
    fit = CCFFit(model, data, device=0)
    eng, _ = fit._fit_engine({})
    dev = torch.device("cuda", 0)
    allrows = params_to_rows(synthetic_batch(65536))
    for n in (1, 8, 30, 64, 128, 256, 400, 600, 700, 900, 1024, 1500, 2048, 3000, 4096, 6000, 8192, 16384):
        d_params = torch.from_numpy(allrows[:n].copy()).to(dev)
        d_theory = torch.empty((n, P), dtype=torch.float64, device=dev)
        d_chi2 = torch.empty(n, dtype=torch.float64, device=dev)
        d_lnl = torch.empty(n, dtype=torch.float64, device=dev)
        out = []
        old_rule = 1 if n >= 888 else min(30, -(-888 // n))
        for label, ns_opt in (("auto", 0), ("old", old_rule), ("1", 1), ("2", 2), ("3", 3), ("5", 5), ("10", 10), ("30", 30)):
            eng.set_option("nsplit", ns_opt)
            ts = []
            for _ in range(12):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                eng.likelihood_ptr(d_params.data_ptr(), n, d_theory.data_ptr(), d_chi2.data_ptr(), d_lnl.data_ptr())
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            out.append(f"{label}:{1e3 * float(np.median(ts[2:])):8.1f}")
        print(f"n={n:6d}  us  " + "  ".join(out))
    fit.close()


if __name__ == "__main__":
    main()
