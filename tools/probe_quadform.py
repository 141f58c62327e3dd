#!/usr/bin/env python
"""GPU: the 60 x 60 quadratic forms of K2 as warp-tiled FMA against FP64 DMMA (vb200p_quad_probe), on the last
precision matrix of the BOSS covariance stack and 65,536 random residual vectors.  Prints time, achieved TFLOP/s
(2 p^2 + 2 p flop per row, p = 60) and the largest difference between the two arrangements and numpy."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import boss_blocks  # noqa: E402
from victor_b200 import CCFFit, _probes  # noqa: E402

fit = CCFFit(*boss_blocks())
p = 60
P = np.zeros((64, 64))
P[:p, :p] = fit.icov[-1]
rng = np.random.default_rng(5)
n = 65536
R = np.zeros((n, 64))
R[:, :p] = rng.standard_normal((n, p)) * np.sqrt(np.diag(fit.covmat[-1]))
want = np.einsum("ni,ij,nj->n", R, P, R)
lib = _probes.load()
flop = n * (2 * p * p + 2 * p)
res = {}
for kind, name in ((0, "warp-tiled FMA"), (1, "FP64 DMMA m8n8k4")):
    q = np.empty(n)
    ms = ctypes.c_double()
    rc = lib.vb200p_quad_probe(0, kind, R.ctypes.data, P.ctypes.data, n, 20, q.ctypes.data, ctypes.byref(ms))
    assert rc == 0, _probes.last_error()
    res[kind] = q
    print(f"{name:18s} {ms.value * 1e3:8.1f} us per 65,536 rows   {flop / (ms.value * 1e-3) / 1e12:6.2f} TFLOP/s (algorithmic)   "
          f"max |q - numpy| / q = {np.max(np.abs(q - want) / np.abs(want)):.2e}")
print(f"max relative difference FMA vs DMMA: {np.max(np.abs(res[0] - res[1]) / np.abs(want)):.2e}")
print("K2 in the bench step: 0.57 ms per 65,536 rows (both brackets, data-vector PCHIP, log-det), 1.9 % of 28.5 ms")
