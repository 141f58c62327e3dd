#!/usr/bin/env python
"""GPU: shared-memory crossbar against the constant-bank path for per-lane table look-ups (vb200p_load_probe).
Prints cycles per warp-cubic per SM sub-partition at 1965 MHz for 1, 2, 3, 4, 8, 32 distinct cells per warp."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from victor_b200 import _probes  # noqa: E402

lib = _probes.load()
ITERS, BPS = 4096, 4
for path, name in ((0, "LDS.128 x2"), (1, "LDC.64 x4 ")):
    for distinct in (1, 2, 3, 4, 8, 32):
        for stride in (0, 5):
            ms = ctypes.c_double()
            rc = lib.vb200p_load_probe(0, path, distinct, stride, BPS, ITERS, ctypes.byref(ms))
            assert rc == 0, _probes.last_error()
            warps_per_smsp = BPS * 8 / 4
            cubics = ITERS * 4 * warps_per_smsp
            cyc = ms.value * 1e-3 * 1.965e9 / cubics
            print(f"{name} distinct={distinct:2d} stride={stride} ms={ms.value:8.3f} cycles per warp-cubic per SMSP = {cyc:6.2f}")
