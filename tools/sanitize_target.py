#!/usr/bin/env python
"""Small run through every kernel variant, meant to be executed under compute-sanitizer memcheck."""
import copy
import os
import sys

import numpy as np
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from victor_b200 import CCFFit, CCFModel  # noqa: E402

with open(os.path.join(ROOT, "config", "boss_config.yaml")) as fh:
    info = yaml.full_load(fh)
info["model"]["dir"] = info["data"]["dir"] = ROOT
fit = CCFFit(copy.deepcopy(info["model"]), copy.deepcopy(info["data"]), device=0)
P = np.array([[0.47, 0.37, 380.0, 1.0, 1.0], [0.6, 0.31, 250.0, 1.04, 0.96], [1.1, 0.7, 60.0, 0.8, 1.2]])
print("streaming", fit.log_likelihood_batch(P)[1])
for kw in ({"rsd_model": "dispersion"}, {"rsd_model": "kaiser"}, {"assume_isotropic": False},
           {"velocity_nodes": 100, "mu_nodes": 200}):
    print(kw, fit.log_likelihood_batch(P[:2], **kw)[1] if "mu_nodes" not in kw else
          fit.theory_multipole_vector_batch(fit.s, P[:2], poles=[0, 2, 4], **kw)[:, :2])
eng, _ = fit._fit_engine({})
for opt in ({"fast_math": 0}, {"ilp": 1}, {"newton": 2}, {"nsplit": 7}):
    for k, v in opt.items():
        eng.set_option(k, v)
    print(opt, fit.log_likelihood_batch(P)[1])
for k, v in (("fast_math", 1), ("ilp", 0), ("newton", 3), ("nsplit", 0)):
    eng.set_option(k, v)
# batch mode (one block per row): separate K2, then the fused epilogue in the tuned and the general kernels
rng = np.random.default_rng(4)
n = 1024
B = np.column_stack([rng.uniform(0.05, 1.5, n), rng.uniform(0.2, 0.6, n), rng.uniform(100, 500, n),
                     rng.uniform(0.9, 1.1, n), rng.uniform(0.9, 1.1, n)])
for fuse in (0, 2):
    for kw in ({}, {"rsd_model": "dispersion"}, {"rsd_model": "kaiser"}):
        e2, _ = fit._fit_engine(kw)
        e2.set_option("fuse", fuse)
        print("batch fuse", fuse, kw, fit.log_likelihood_batch(B, **kw)[1][:2])
        e2.set_option("fuse", 1)
# MCMC-sized calls: graph replay, block-per-row K2; pairwise points (theory_xi_2D)
for i in range(3):
    print("n=1", fit.log_likelihood({"fsigma8": 0.47 + 0.01 * i, "beta": 0.37, "sigma_v": 380.0, "epsilon": 1.0}))
print("n=5", fit.log_likelihood_batch(P[[0, 1, 2, 0, 1]])[1])
xi2d = fit.theory_xi_2D({"fsigma8": 0.47, "beta": 0.37, "sigma_v": 380.0, "epsilon": 1.0}, rmax=60)
print("xi2d", xi2d(10.0, 20.0), fit.theory_xi_pairs_batch(np.linspace(1, 100, 333), np.linspace(-1, 1, 333), P,
                                                           rsd_model="dispersion")[:, :2])
fit.close()
me = copy.deepcopy(info["model"])
me["velocity_pdf"]["mean"]["empirical_corr"] = True
fe = CCFFit(me, copy.deepcopy(info["data"]), device=0)
print("empirical", fe.log_likelihood_batch({"fsigma8": P[:, 0], "beta": P[:, 1], "sigma_v": P[:, 2], "Av": [0.0, 0.5, -0.5]})[1])
fe.close()
m = copy.deepcopy(info["model"])
m["input_model_data_file"] = "tests/golden/model_sv2d_inputs.npz"
m["velocity_pdf"]["dispersion"]["template_keys"] = ["rsv", "musv", "sigmav2d"]
f2 = CCFFit(m, copy.deepcopy(info["data"]), device=0)
print("sv2d", f2.log_likelihood_batch(P[:2])[1])
f2.close()
with open(os.path.join(ROOT, "config", "example_model_input.yaml")) as fh:
    em = yaml.full_load(fh)["model"]
em["dir"] = ROOT
ex = CCFModel(em, device=0)
print("example", ex.theory_multipole_vector_batch(np.linspace(0.01, 3, 40), {"fsigma8": 0.47, "sigma_v": 7.0, "epsilon": 1.0},
                                                  poles=[0, 2, 4])[0, :3])
ex.close()
print("done")
