#!/usr/bin/env python
"""A few n = 1 likelihood calls with device-resident buffers, first through the one-launch kernel (k_small), then
through K1 + K2: run under `ncu --metrics gpu__time_duration.sum` to read the kernels' own durations."""
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
for tiny in (1, 0):
    eng.set_option("tiny", tiny)
    for _ in range(6):
        eng.likelihood_ptr(rows.data_ptr(), 1, None, out[1].data_ptr(), out[0].data_ptr(), None)
        torch.cuda.synchronize()
    print("tiny", tiny, out.cpu().numpy().ravel())
fit.close()
