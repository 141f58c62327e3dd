#!/usr/bin/env python
"""GPU: warm duration of an n = 1 likelihood evaluation on the device, without host gaps: a CUDA graph holding 100
back-to-back device-buffer calls is replayed, so successive calls queue behind each other on the stream and the
per-call time is the kernels' own (launch overhead inside a graph is ~1 us).  k_small (one launch) against K1 + K2."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import boss_blocks  # noqa: E402
from victor_b200 import CCFFit  # noqa: E402
from victor_b200.model import params_to_rows  # noqa: E402

fit = CCFFit(*boss_blocks(), device=0)
prm = {"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380.0, "epsilon": 1.0}
CALLS = 100
for kw in ({}, {"rsd_model": "dispersion"}, {"assume_isotropic": False}):
    eng, _ = fit._fit_engine(kw)
    for n in (1, 2, 3, 4):
        rows = torch.from_numpy(np.repeat(params_to_rows(dict(prm)), n, axis=0)).cuda()
        out = torch.empty((2, n), dtype=torch.float64, device="cuda")
        for tiny in (1, 0):
            eng.set_option("tiny", tiny)
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for _ in range(3):
                    eng.likelihood_ptr(rows.data_ptr(), n, None, out[1].data_ptr(), out[0].data_ptr(), stream.cuda_stream)
                stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream):
                    for _ in range(CALLS):
                        eng.likelihood_ptr(rows.data_ptr(), n, None, out[1].data_ptr(), out[0].data_ptr(),
                                           torch.cuda.current_stream().cuda_stream)
                ts = []
                for _ in range(12):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    g.replay()
                    b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b) * 1e3 / CALLS)
            print(f"{kw or 'streaming'} n={n} tiny={tiny}: {np.median(ts[2:]):7.2f} us per call on the device  "
                  f"(lnl {float(out[0, 0]):.6f})")
        eng.set_option("tiny", 1)
fit.close()
