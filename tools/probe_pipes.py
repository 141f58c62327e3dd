#!/usr/bin/env python
"""GPU: which pipe do FP64 conversions use (throughput alone vs interleaved with DFMA), and how
accurate are the MUFU rsqrt / rcp seeds?  Prints a small JSON report."""
import ctypes
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from victor_b200 import _probes as _lib  # noqa: E402

lib = _lib.load()
names = {0: "dfma", 1: "f2i_f64_floor", 2: "dfma+f2i", 3: "f2f_f32_f64", 4: "dfma+f2f", 5: "i2f_f64", 6: "dfma+i2f"}
rep = {}
for mode, name in names.items():
    ms = ctypes.c_double()
    rc = lib.vb200p_pipe_probe(0, mode, 2048, ctypes.byref(ms))
    assert rc == 0, _lib.last_error()
    rep[name + "_ms"] = ms.value
rng = np.random.default_rng(0)
x = np.ascontiguousarray(np.concatenate([rng.uniform(0.25, 4.0, 400000), 10.0 ** rng.uniform(-3, 5, 100000)]))
out = np.empty(2 * len(x))
assert lib.vb200p_seed_probe(0, x.ctypes.data, len(x), out.ctypes.data) == 0
rs, rc_ = out[:len(x)], out[len(x):]
e_rs = rs * np.sqrt(x) - 1
e_rc = rc_ * x - 1
rep["rsqrt_seed_max_rel"] = float(np.abs(e_rs).max())
rep["rsqrt_seed_log2"] = float(np.log2(np.abs(e_rs).max()))
rep["rcp_seed_max_rel"] = float(np.abs(e_rc).max())
rep["rcp_seed_log2"] = float(np.log2(np.abs(e_rc).max()))
rep["rsqrt_seed_mean_rel"] = float(e_rs.mean())
rep["rcp_seed_mean_rel"] = float(e_rc.mean())
print(json.dumps(rep, indent=1))
