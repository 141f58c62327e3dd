"""TEST INFRASTRUCTURE ONLY - import shims that let the UNMODIFIED reference run here.

Used by ``oracle/make_golden.py`` (dev container only; the GPU box has no /root/reference)
to produce the committed fixtures in ``tests/golden/``.  Nothing in the product imports it.

The reference (victor 0.1.4) imports five things this image does not have or that the
installed scipy (1.18) has removed.  Each shim reproduces the documented behaviour of the
missing piece and touches no reference code:

1. ``h5py.File``                       -> victor_b200.io_hdf5 (dict of whole datasets)
2. ``matplotlib[.pyplot|.colors|.cm]`` -> attribute-permissive dummies (plotting unused)
3. ``astropy.cosmology.LambdaCDM``     -> closed-form E(z), no radiation (Tcmb0 = 0 default)
4. ``scipy.integrate.simps``           -> ``scipy.integrate.simpson`` (same call signature
                                          as used at victor/ccf_model.py:690)
5. ``scipy.interpolate.interp2d``      -> RectBivariateSpline(kx=ky=k, s=0) on the regular
                                          grid, which is the FITPACK ``regrid`` fit legacy
                                          interp2d used for gridded input
6. ``cobaya.likelihood.Likelihood``    -> stand-in base class (class attributes from the
                                          yaml defaults, then ``initialize()``)
"""
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


class _Permissive(types.ModuleType):
    """Module whose every attribute is another permissive callable object."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _PermissiveObj()
        setattr(self, name, obj)
        return obj


class _PermissiveObj:
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _PermissiveObj()

    def __call__(self, *a, **k):
        return _PermissiveObj()

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return _PermissiveObj()


class _H5Dataset:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, key):
        return self._arr[key]


class _H5File:
    def __init__(self, fn, mode="r"):
        from victor_b200.io_hdf5 import _read_native
        self._data = _read_native(fn)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def keys(self):
        return self._data.keys()

    def __getitem__(self, key):
        return _H5Dataset(self._data[key])


class _Quantity:
    def __init__(self, value):
        self.value = value


class _LambdaCDM:
    def __init__(self, H0, Om0, Ode0, **kw):
        self.H0, self.Om0, self.Ode0 = H0, Om0, Ode0
        self.Ok0 = 1.0 - Om0 - Ode0

    def efunc(self, z):
        zp1 = 1.0 + np.asarray(z, dtype=float)
        return np.sqrt(self.Om0 * zp1 ** 3 + self.Ok0 * zp1 ** 2 + self.Ode0)

    def H(self, z):
        return _Quantity(self.H0 * self.efunc(z))

    def Om(self, z):
        zp1 = 1.0 + np.asarray(z, dtype=float)
        return self.Om0 * zp1 ** 3 / self.efunc(z) ** 2


class _Interp2d:
    """Regular-grid stand-in for the removed scipy.interpolate.interp2d."""

    def __init__(self, x, y, z, kind="linear", **kw):
        from scipy.interpolate import RectBivariateSpline
        k = {"linear": 1, "cubic": 3, "quintic": 5}[kind]
        x = np.asarray(x, dtype=float).ravel()
        y = np.asarray(y, dtype=float).ravel()
        z = np.asarray(z, dtype=float)
        if z.shape != (len(y), len(x)):
            raise ValueError("interp2d shim: z must have shape (len(y), len(x))")
        self._lim = (x.min(), x.max(), y.min(), y.max())
        self._spl = RectBivariateSpline(x, y, z.T, kx=k, ky=k, s=0)

    def __call__(self, x, y):
        x = np.clip(np.sort(np.atleast_1d(np.asarray(x, dtype=float))), self._lim[0], self._lim[1])
        y = np.clip(np.sort(np.atleast_1d(np.asarray(y, dtype=float))), self._lim[2], self._lim[3])
        out = self._spl(x, y).T  # (len(y), len(x)) like legacy interp2d
        if out.shape[1] == 1 and out.shape[0] == 1:
            return out.ravel()
        if out.shape[1] == 1:
            # legacy interp2d squeezed to 1-D when one argument was scalar; the reference does
            # ``f(r_j, mu).T[0]`` (victor/utils.py:55), which works for either (N,1) or (N,)
            # only when 2-D, so keep the 2-D (N,1) form.
            return out
        return out


class _Likelihood:
    """Stand-in for cobaya.likelihood.Likelihood: yaml defaults + overrides, then initialize()."""

    def __init__(self, info=None, **kw):
        import yaml
        import inspect
        cls_file = inspect.getfile(type(self))
        yml = os.path.splitext(cls_file)[0] + ".yaml"
        defaults = {}
        if os.path.isfile(yml):
            with open(yml) as fh:
                defaults = yaml.full_load(fh) or {}
        defaults.update(info or {})
        defaults.update(kw)
        for key, val in defaults.items():
            setattr(self, key, val)
        self.initialize()

    def initialize(self):
        pass


def find_reference():
    """Where an importable copy of the unmodified reference lives, or None: ``baseline/_ref`` (the offline
    ``pip install --no-deps --target baseline/_ref`` of /root/reference; git-ignored, shipped to the GPU box)
    first, then the read-only checkout in the dev container."""
    if os.environ.get("VB200_NO_REFERENCE"):      # tests: exercise the fallback to the oracle port
        return None
    for root in (os.path.join(_ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(root, "victor", "ccf_fit.py")):
            return root
    return None


def npy_twins(blocks, outdir):
    """Rewrite the (model, data) option blocks of this repository so that the UNMODIFIED reference can read
    their input files: the committed ``.npz`` re-encodings of the reference's HDF5 inputs become ``.npy``
    dict files in ``outdir`` -- a format the reference reads itself (ccf_model.py:62-63, ccf_fit.py:51-52),
    so no file-reader shim sits on its path."""
    import copy
    model, data = copy.deepcopy(blocks[0]), copy.deepcopy(blocks[1])

    def twin(base_dir, rel):
        src = os.path.join(base_dir, rel)
        dst = os.path.join(outdir, os.path.splitext(os.path.basename(rel))[0] + ".npy")
        if not os.path.isfile(dst):
            with np.load(src) as z:
                np.save(dst, {k: z[k] for k in z.files}, allow_pickle=True)
        return os.path.basename(dst)

    model["input_model_data_file"] = twin(model.get("dir", ""), model["input_model_data_file"])
    model["dir"] = outdir
    for key in ("redshift_space_ccf", "covariance_matrix"):
        data[key]["data_file"] = twin(data.get("dir", ""), data[key]["data_file"])
    data["dir"] = outdir
    return model, data


def install(reference_root=None):
    """Install the shims and put the reference on sys.path.  Idempotent."""
    reference_root = reference_root or find_reference()
    if not reference_root or not os.path.isdir(os.path.join(reference_root, "victor")):
        raise FileNotFoundError(f"reference not found under {reference_root}")

    h5 = types.ModuleType("h5py")
    h5.File = _H5File
    sys.modules.setdefault("h5py", h5)

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.cm"):
        sys.modules.setdefault(name, _Permissive(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]

    astropy = types.ModuleType("astropy")
    cosmology = types.ModuleType("astropy.cosmology")
    cosmology.LambdaCDM = _LambdaCDM
    astropy.cosmology = cosmology
    sys.modules.setdefault("astropy", astropy)
    sys.modules.setdefault("astropy.cosmology", cosmology)

    import scipy.integrate
    import scipy.interpolate
    if not hasattr(scipy.integrate, "simps"):
        scipy.integrate.simps = scipy.integrate.simpson
    if not hasattr(scipy.interpolate, "interp2d") or scipy.interpolate.interp2d is not _Interp2d:
        scipy.interpolate.interp2d = _Interp2d
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid

    cobaya = types.ModuleType("cobaya")
    cl = types.ModuleType("cobaya.likelihood")
    cl.Likelihood = _Likelihood
    cobaya.likelihood = cl
    sys.modules.setdefault("cobaya", cobaya)
    sys.modules.setdefault("cobaya.likelihood", cl)

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    import warnings
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import victor  # noqa: F401  (the unmodified reference)
    return victor
