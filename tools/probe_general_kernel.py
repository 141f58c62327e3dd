#!/usr/bin/env python
"""GPU: throughput of the configurations that run on the GENERAL kernel (k1_general.cuh): kaiser / euclid_special,
real-space ccf measured from data (anisotropic, MD covariance), sigma_v(r, mu) templates -- 65,536 rows each,
device-resident buffers, CUDA events.  The tuned-kernel models are printed beside them for scale."""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import boss_blocks, synthetic_batch  # noqa: E402
from victor_b200 import CCFFit  # noqa: E402
from victor_b200.model import params_to_rows  # noqa: E402


def measured_blocks():
    model, data = boss_blocks()
    model["input_model_data_file"] = "data/boss_dr12_cmass/cmass_measured_model.npz"
    model["realspace_ccf"]["from_data"] = True
    model["realspace_ccf"]["assume_isotropic"] = False
    data["covariance_matrix"]["data_file"] = "data/boss_dr12_cmass/cmass_variable_anisotropic_MD_covariance.npz"
    return model, data


def sv2d_blocks():
    model, data = boss_blocks()
    model["input_model_data_file"] = "tests/golden/model_sv2d_inputs.npz"
    model["velocity_pdf"]["dispersion"]["template_keys"] = ["rsv", "musv", "sigmav2d"]
    return model, data


n = 65536
rows = params_to_rows(synthetic_batch(n))
d_rows = torch.from_numpy(rows).cuda()
out = torch.empty((2, n), dtype=torch.float64, device="cuda")
cases = [("BOSS", boss_blocks, {}), ("BOSS", boss_blocks, {"rsd_model": "dispersion"}), ("BOSS", boss_blocks, {"assume_isotropic": False}),
         ("BOSS", boss_blocks, {"rsd_model": "kaiser"}), ("BOSS", boss_blocks, {"rsd_model": "euclid_special"}),
         ("measured model (from data, anisotropic)", measured_blocks, {}),
         ("measured model (from data, anisotropic)", measured_blocks, {"rsd_model": "dispersion"}),
         ("sigma_v(r, mu) template", sv2d_blocks, {}), ("sigma_v(r, mu) template", sv2d_blocks, {"rsd_model": "dispersion"})]
fits = {}
for name, blocks, kw in cases:
    if name not in fits:
        fits[name] = CCFFit(*blocks(), device=0)
    eng, _ = fits[name]._fit_engine(kw)
    eng.set_option("tuned", int(os.environ.get("VB200_TUNED", "1")))
    ts = []
    for _ in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.likelihood_ptr(d_rows.data_ptr(), n, None, out[1].data_ptr(), out[0].data_ptr(), None)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    nonfinite = int(np.count_nonzero(~np.isfinite(out[0].cpu().numpy())))
    print(f"{name:42s} {str(kw or 'streaming'):38s} {min(ts):8.2f} ms  {n / (min(ts) * 1e-3):10.4g} evals/s  non-finite rows {nonfinite}  chi2[0] {float(out[1, 0]):.9f}")
for f in fits.values():
    f.close()
