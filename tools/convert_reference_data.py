#!/usr/bin/env python
"""Convert the HDF5 inputs shipped with the reference into .npz fixtures under data/.

The GPU box has no /root/reference, so the input tables the BASELINE configurations name
(BOSS DR12 CMASS void-galaxy model / data / covariance files and the example void model)
travel with this repository as plain ``numpy.savez`` archives holding exactly the datasets
of the HDF5 originals (same keys, same float64 values, bit for bit).

Run in the dev container:   python tools/convert_reference_data.py [/root/reference]
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from victor_b200.io_hdf5 import read_hdf5  # noqa: E402

SAMPLE = "zobovVoids_reconRs10_0.43z0.7_medianRvcut"
BOSS = "data/BOSS_DR12_CMASS_data"
FILES = {
    f"{BOSS}/CMASS_{SAMPLE}_PatchyMean_model.hdf5": "data/boss_dr12_cmass/cmass_patchymean_model.npz",
    f"{BOSS}/CMASS_{SAMPLE}_measured_model.hdf5": "data/boss_dr12_cmass/cmass_measured_model.npz",
    f"{BOSS}/CMASS_{SAMPLE}_data.hdf5": "data/boss_dr12_cmass/cmass_data.npz",
    f"{BOSS}/CMASS_{SAMPLE}_fixed_D_covariance.hdf5": "data/boss_dr12_cmass/cmass_fixed_D_covariance.npz",
    f"{BOSS}/CMASS_{SAMPLE}_variable_D_covariance.hdf5": "data/boss_dr12_cmass/cmass_variable_D_covariance.npz",
    f"{BOSS}/CMASS_{SAMPLE}_variable_isotropic_MD_covariance.hdf5":
        "data/boss_dr12_cmass/cmass_variable_isotropic_MD_covariance.npz",
    f"{BOSS}/CMASS_{SAMPLE}_variable_anisotropic_MD_covariance.hdf5":
        "data/boss_dr12_cmass/cmass_variable_anisotropic_MD_covariance.npz",
    "data/example_data/example_void_model.hdf5": "data/example/example_void_model.npz",
}


def main(ref_root):
    for src, dst in FILES.items():
        arrays = read_hdf5(os.path.join(ref_root, src))
        out = os.path.join(ROOT, dst)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        np.savez(out, **arrays)
        back = np.load(out)
        assert all(np.array_equal(back[k], v) for k, v in arrays.items())
        print(f"{src} -> {dst}: " + ", ".join(f"{k}{v.shape}" for k, v in arrays.items()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
