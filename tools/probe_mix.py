#!/usr/bin/env python
"""GPU: DFMA issue model.  Cycles per DFMA warp-instruction per SM sub-partition as a function of
independent chains per thread, resident warps, and interleaved non-FP64 instructions."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from victor_b200 import _probes as _lib  # noqa: E402

lib = _lib.load()
CLK = 1.965e9
ITERS = 4096
print("chains mix kind warps/SMSP   ms   cycles per DFMA per SMSP  (2.0 = nominal peak)")
for chains, mix, kind in [(1, 0, 0), (2, 0, 0), (4, 0, 0), (8, 0, 0), (2, 1, 0), (4, 1, 0), (8, 1, 0), (4, 2, 0),
                          (2, 1, 1), (4, 1, 1), (8, 1, 1), (4, 0, 2), (8, 0, 2), (4, 0, 3), (4, 0, 4),
                          (4, 0, 5), (8, 0, 5), (4, 0, 6), (8, 0, 6), (4, 0, 7), (8, 0, 7)][int(os.environ.get("PROBE_FROM", 0)):]:
    for bps in (1, 2, 4, 8, 16):
        ms = ctypes.c_double()
        rc = lib.vb200p_mix_probe(0, chains, mix, kind, bps, ITERS, ctypes.byref(ms))
        assert rc == 0, _lib.last_error()
        per_iter = 12 * chains if kind in (3, 4) else 8 * chains
        per = ms.value * 1e-3 * CLK / (ITERS * per_iter * bps)
        print(f"{chains:6d} {mix:3d} {kind:4d} {bps:10d} {ms.value:8.3f} {per:8.2f}")
