#!/usr/bin/env python
"""Where an n = 1 likelihood call (MCMC step) spends its time: the C-ABI call alone, the engine wrapper,
CCFFit.log_likelihood, CCFLikelihood.calculate.  Median microseconds over many calls."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import boss_blocks  # noqa: E402
from victor_b200.likelihoods import CCFLikelihood  # noqa: E402
from victor_b200.model import params_to_rows  # noqa: E402


def med(fn, n=3000, warm=200):
    for _ in range(warm):
        fn()
    t = np.empty(n)
    for i in range(n):
        t0 = time.perf_counter()
        fn()
        t[i] = time.perf_counter() - t0
    return float(np.median(t) * 1e6), float(np.percentile(t, 95) * 1e6)


def main():
    model, data = boss_blocks()
    like = CCFLikelihood({"model": model, "data": data, "device": 0})
    fit = like.ccf
    prm = {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380.0, "epsilon": 1.0, "alpha": 1}
    eng, _ = fit._fit_engine({})
    rows = params_to_rows(dict(prm))
    chi2, lnl = np.empty(1), np.empty(1)
    lib, h = eng.lib, eng.handle
    rp, cp, lp = rows.ctypes.data, chi2.ctypes.data, lnl.ctypes.data
    print("C ABI call (ctypes, prebuilt buffers)  us median/p95:", med(lambda: lib.vb200_likelihood(h, rp, 1, None, cp, lp, None)))
    print("Engine.likelihood(rows)                us median/p95:", med(lambda: eng.likelihood(rows)))
    print("params_to_rows(dict)                   us median/p95:", med(lambda: params_to_rows(prm)))
    print("CCFFit.log_likelihood(dict)            us median/p95:", med(lambda: fit.log_likelihood(prm)))
    st = {}
    print("CCFLikelihood.calculate(**dict)        us median/p95:", med(lambda: like.calculate(st, **prm)))
    for opt, val in (("mapped", 0), ("tiny", 0), ("graph", 0), ("tiny", 1), ("graph", 1), ("mapped", 1)):
        eng.set_option(opt, val)
        print(f"C ABI call after {opt}={val}              us median/p95:", med(lambda: lib.vb200_likelihood(h, rp, 1, None, cp, lp, None)))
    # device-resident buffers: kernel time of the one-launch path by CUDA events
    import torch
    d_rows = torch.from_numpy(rows).cuda()
    d_out = torch.empty((2, 1), dtype=torch.float64, device="cuda")
    for tiny in (1, 0):
        eng.set_option("tiny", tiny)
        ts = []
        for i in range(300):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.likelihood_ptr(d_rows.data_ptr(), 1, None, d_out[1].data_ptr(), d_out[0].data_ptr(), None)
            b.record()
            torch.cuda.synchronize()
            if i >= 50:
                ts.append(a.elapsed_time(b) * 1e3)
        print(f"CUDA events around the device-buffer call, tiny={tiny}: us median {np.median(ts):.2f}")
    eng.set_option("tiny", 1)
    fit.close()


if __name__ == "__main__" and "--device" not in sys.argv:
    main()


def device_side():
    """The same call with device-resident buffers: K1 alone, K1 + K2, and CUDA-event kernel times."""
    import torch
    model, data = boss_blocks()
    from victor_b200 import CCFFit
    fit = CCFFit(model, data, device=0)
    eng, _ = fit._fit_engine({})
    rows = params_to_rows({"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380.0, "epsilon": 1.0, "alpha": 1})
    dev = torch.device("cuda", 0)
    d_rows = torch.from_numpy(rows).to(dev)
    d_th = torch.empty(60, dtype=torch.float64, device=dev)
    d_c2 = torch.empty(1, dtype=torch.float64, device=dev)
    d_ll = torch.empty(1, dtype=torch.float64, device=dev)

    def k1_only():
        eng.likelihood_ptr(d_rows.data_ptr(), 1, d_th.data_ptr(), None, None)
        torch.cuda.synchronize()

    def k1_k2():
        eng.likelihood_ptr(d_rows.data_ptr(), 1, d_th.data_ptr(), d_c2.data_ptr(), d_ll.data_ptr())
        torch.cuda.synchronize()

    print("device buffers, K1 only + sync         us median/p95:", med(k1_only))
    print("device buffers, K1 + K2 + sync         us median/p95:", med(k1_k2))
    print("empty torch.cuda.synchronize()         us median/p95:", med(torch.cuda.synchronize))
    for name, fn in (("K1", lambda: eng.likelihood_ptr(d_rows.data_ptr(), 1, d_th.data_ptr(), None, None)),
                     ("K1+K2", lambda: eng.likelihood_ptr(d_rows.data_ptr(), 1, d_th.data_ptr(), d_c2.data_ptr(), d_ll.data_ptr()))):
        ts = []
        for _ in range(300):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f"CUDA events around {name:6s}              us median:", float(np.median(ts)))
    fit.close()


if __name__ == "__main__" and "--device" in sys.argv:
    device_side()
