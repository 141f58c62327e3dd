#!/usr/bin/env python
"""GPU, diagnostic build (-DVB200_SMALL_TIMING, VICTOR_B200_LIB=build/libvb_timing.so): globaltimer stamps inside
k_small -- where an n = 1 evaluation spends its time on the device."""
import ctypes
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
eng, _ = fit._fit_engine({})
rows = torch.from_numpy(params_to_rows({"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380.0, "epsilon": 1.0})).cuda()
out = torch.empty((2, 1), dtype=torch.float64, device="cuda")
names = ["start (block 0)", "scalars + barrier", "cell records + barrier", "quadrature + butterfly + barrier",
         "ticket (block 0)", "last block: is last", "last: projection done", "last: quadratic forms done", "last: end"]
acc = []
for it in range(40):
    eng.likelihood_ptr(rows.data_ptr(), 1, None, out[1].data_ptr(), out[0].data_ptr(), None)
    torch.cuda.synchronize()
    st = np.zeros(16, dtype=np.uint64)
    assert eng.lib.vb200_debug_stamps(st.ctypes.data_as(ctypes.c_void_p)) == 0
    if it >= 10:
        acc.append(st[:9].astype(np.int64) - int(st[0]))
acc = np.median(np.array(acc), axis=0)
for n, t in zip(names, acc):
    print(f"{t / 1e3:8.2f} us  {n}")
fit.close()
