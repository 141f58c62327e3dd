"""Small host-side helpers shared by the drop-in classes.

``InputError`` keeps the name and role of ``victor.utils.InputError`` (victor/utils.py:5):
it is raised for anything wrong with user input (missing files or keys, bad shapes,
invalid option values).
"""
import logging
import os

import numpy as np

# the reference reports failures with bare print() (ccf_fit.py:402, 408, 449, 478-479); here they go through
# the standard logging module: logging.getLogger("victor_b200").setLevel(...) to see or silence them
log = logging.getLogger("victor_b200")

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


class GridInterpolator2D:
    """Stand-in for the ``scipy.interpolate.interp2d`` objects the reference returns (removed in
    scipy 1.14): spline of degree 1 / 3 / 5 through ``z[len(y)][len(x)]`` on the regular grid (x, y).

    Called as ``f(x, y)`` it follows legacy interp2d: arguments are sorted, arguments outside the grid
    are moved to its edge, the result has shape (len(y), len(x)) and loses its first axis when
    ``y`` is a single value.
    """

    def __init__(self, x, y, z, kind="linear"):
        from scipy.interpolate import RectBivariateSpline
        degree = {"linear": 1, "cubic": 3, "quintic": 5}.get(kind)
        if degree is None:
            raise ValueError(f"Unsupported interpolation type {kind}")
        self.x = np.asarray(x, dtype=np.float64).ravel()
        self.y = np.asarray(y, dtype=np.float64).ravel()
        z = np.asarray(z, dtype=np.float64)
        if z.shape != (len(self.y), len(self.x)):
            raise ValueError(f"z must have shape ({len(self.y)}, {len(self.x)}), got {z.shape}")
        self._spline = RectBivariateSpline(self.x, self.y, z.T, kx=degree, ky=degree, s=0)

    def __call__(self, x, y):
        xq = np.clip(np.sort(np.atleast_1d(np.asarray(x, dtype=np.float64)).ravel()), self.x[0], self.x[-1])
        yq = np.clip(np.sort(np.atleast_1d(np.asarray(y, dtype=np.float64)).ravel()), self.y[0], self.y[-1])
        out = self._spline(xq, yq).T
        return out[0] if out.shape[0] == 1 else out


def multipoles_from_fn(frmu, r, ell=(0, 2, 4), even=True, npts=200):
    """Legendre multipoles of ``frmu(r, mu)`` at the radii ``r`` (reference: victor/utils.py:9-58):
    trapezoid over ``npts`` mu values, on [0, 1] doubled for a function even in mu, else on [-1, 1].
    Returns ``{'l': array}``."""
    from scipy.special import legendre
    ells = np.atleast_1d(ell)
    mu = np.linspace(0.0, 1.0, npts) if even else np.linspace(-1, 1, npts)
    out = {f"{l}": np.zeros(len(r)) for l in ells}
    for l in ells:
        weight = (2 * l + 1 if even else (2 * l + 1) / 2) * legendre(l)(mu)
        for j, rj in enumerate(r):
            out[f"{l}"][j] = trapezoid(np.asarray(frmu(rj, mu)).reshape(len(mu)) * weight, mu)
    return out


def fn_from_multipoles(r, poles, multipoles, npts=200):
    """f(r, mu) = sum_l f_l(r) L_l(mu) as an interpolating function (reference: victor/utils.py:60-94)."""
    from scipy.special import legendre
    poles = [poles] if isinstance(poles, int) else poles
    multipoles = np.asarray(multipoles)
    if multipoles.shape != (len(poles), len(r)):
        raise ValueError(f"Wrong shape of multipoles: expected ({len(poles)}, {len(r)}), but received "
                         f"{multipoles.shape}")
    mu = np.linspace(-1, 1, npts)
    grid = np.zeros((len(mu), len(r)))
    for i, l in enumerate(poles):
        grid += legendre(l)(mu)[:, None] * multipoles[i]
    return GridInterpolator2D(r, mu, grid)
