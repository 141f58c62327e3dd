"""Small host-side helpers shared by the drop-in classes.

``InputError`` keeps the name and role of ``victor.utils.InputError`` (victor/utils.py:5):
it is raised for anything wrong with user input (missing files or keys, bad shapes,
invalid option values).
"""
import os

import numpy as np

from .io_hdf5 import read_hdf5


class InputError(Exception):
    """Error raised when something is wrong with the input data."""


_HDF5_EXT = (".hdf", ".h4", ".hdf4", ".he2", ".h5", ".hdf5", ".he5", ".h5py")


def load_input_file(path):
    """Read a model / data / covariance input file into ``{key: ndarray}``.

    Formats: ``.npy`` pickled dict and HDF5 as in the reference (victor/ccf_model.py:54-68),
    plus ``.npz`` archives (this repository's fixtures).
    """
    if not os.path.isfile(path):
        raise InputError(f"File {path} not found")
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: np.asarray(z[k]) for k in z.files}
    if path.endswith(".npy"):
        return np.load(path, allow_pickle=True).item()
    if path.endswith(_HDF5_EXT):
        return read_hdf5(path)
    raise InputError(f"Unrecognised input file format: {path}")


def trapezoid(y, x):
    fn = getattr(np, "trapezoid", None) or np.trapz
    return fn(y, x)
