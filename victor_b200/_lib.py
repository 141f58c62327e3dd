"""ctypes binding of libvictor_b200.so (include/victor_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make``; there is no fallback:
a missing library raises ImportError with the build command.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# VICTOR_B200_LIB selects another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("VICTOR_B200_LIB") or os.path.join(_HERE, "libvictor_b200.so")

MAX_POLES = 3
c_double_p = POINTER(c_double)
c_int32_p = POINTER(c_int32)

EXPORTS = ("vb200_version", "vb200_abi_check", "vb200_last_error", "vb200_device_count", "vb200_create", "vb200_destroy",
           "vb200_set_option", "vb200_theory", "vb200_theory_pairs", "vb200_likelihood", "vb200_synchronize",
           "vb200_launch_count")


class ModelTablesC(ctypes.Structure):
    _fields_ = [
        ("iaH", c_double), ("template_sigma8", c_double), ("beta_fixed", c_double), ("inv_h", c_double),
        ("vel_indep_AP", c_int32), ("rsd_model", c_int32), ("n_ell", c_int32),
        ("ells", c_int32 * MAX_POLES), ("beta_dependent", c_int32),
        ("ncell", c_int32), ("nbucket", c_int32), ("maxscan", c_int32),
        ("nbeta", c_int32), ("nx", c_int32), ("nresc", c_int32),
        ("realspace_from_data", c_int32), ("kaiser_approximation", c_int32), ("kaiser_coord_shift", c_int32),
        ("niter", c_int32), ("sv_ny", c_int32), ("vd_beta_dependent", c_int32), ("growth_mode", c_int32),
        ("linear_bias", c_int32),
        ("bias", c_double), ("template_fsigma8", c_double), ("growth_scale", c_double),
        ("origin", c_double_p), ("upper", c_double_p), ("bucket_base", c_int32_p),
        ("beta_grid", c_double_p), ("xi_tab", c_double_p),
        ("v0", c_double_p), ("d0", c_double_p), ("v0b", c_double_p), ("d0b", c_double_p), ("sv", c_double_p),
        ("sv2d", c_double_p), ("sv_ybreaks", c_double_p),
        ("x", c_double_p), ("wx", c_double_p), ("mu_resc", c_double_p), ("w_resc", c_double_p),
    ]


class FitTablesC(ctypes.Structure):
    _fields_ = [
        ("ns", c_int32), ("npoles", c_int32), ("nmu", c_int32), ("data_beta_dependent", c_int32),
        ("nbeta_ccf", c_int32), ("cov_fixed", c_int32), ("nbeta_cov", c_int32),
        ("like_kind", c_int32), ("use_logdet", c_int32),
        ("like_a", c_double), ("like_nm1", c_double),
        ("s", c_double_p), ("mu", c_double_p), ("wmu", c_double_p),
        ("beta_ccf", c_double_p), ("data_tab", c_double_p), ("beta_cov", c_double_p),
        ("icov", c_double_p), ("logdet", c_double_p), ("lam", c_double_p),
    ]


_lib = None


def load():
    """Load the shared library once; raise ImportError (never fall back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()'  or  make -C victor_b200/csrc). "
            "victor_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.vb200_version.restype = c_char_p
    lib.vb200_last_error.restype = c_char_p
    lib.vb200_device_count.restype = c_int
    lib.vb200_abi_check.restype = c_int
    lib.vb200_abi_check.argtypes = [c_int64, c_int64]
    if lib.vb200_abi_check(ctypes.sizeof(ModelTablesC), ctypes.sizeof(FitTablesC)) != 0:
        raise ImportError("libvictor_b200.so does not match this binding (struct sizes differ); rebuild it")
    lib.vb200_create.restype = c_int
    lib.vb200_create.argtypes = [POINTER(ModelTablesC), POINTER(FitTablesC), c_int, POINTER(c_void_p)]
    lib.vb200_destroy.restype = None
    lib.vb200_destroy.argtypes = [c_void_p]
    lib.vb200_set_option.restype = c_int
    lib.vb200_set_option.argtypes = [c_void_p, c_char_p, c_int64]
    lib.vb200_theory.restype = c_int
    lib.vb200_theory.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_int32,
                                 c_void_p, c_int32, c_void_p, c_void_p, c_void_p]
    lib.vb200_theory_pairs.restype = c_int
    lib.vb200_theory_pairs.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]
    lib.vb200_likelihood.restype = c_int
    lib.vb200_likelihood.argtypes = [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.vb200_synchronize.restype = c_int
    lib.vb200_synchronize.argtypes = [c_void_p]
    lib.vb200_launch_count.restype = c_int64
    lib.vb200_launch_count.argtypes = [c_void_p]
    _lib = lib
    return lib


def last_error():
    return load().vb200_last_error().decode("utf-8", "replace")


def _dp(arr):
    return arr.ctypes.data_as(c_double_p)


def pack_model(mt):
    """tables.ModelTables -> (ModelTablesC, keepalive list of contiguous arrays)."""
    keep = {}

    def f64(name, a):
        keep[name] = np.ascontiguousarray(a, dtype=np.float64)
        return _dp(keep[name])

    c = ModelTablesC()
    c.iaH, c.template_sigma8, c.beta_fixed, c.inv_h = mt.iaH, mt.template_sigma8, mt.beta_fixed, mt.inv_h
    c.vel_indep_AP, c.rsd_model, c.n_ell = int(mt.vel_indep_AP), int(mt.rsd_model), int(mt.n_ell)
    for i in range(MAX_POLES):
        c.ells[i] = int(mt.ells[i]) if i < len(mt.ells) else 0
    c.beta_dependent = int(mt.beta_dependent)
    c.ncell, c.nbucket, c.maxscan = int(mt.ncell), len(mt.bucket_base), int(mt.maxscan)
    c.nbeta, c.nx, c.nresc = len(mt.beta_grid), len(mt.x), len(mt.mu_resc)
    c.realspace_from_data, c.kaiser_approximation = int(mt.from_data), int(mt.kaiser_approximation)
    c.kaiser_coord_shift, c.niter = int(mt.kaiser_coord_shift), int(mt.niter)
    c.origin, c.upper = f64("origin", mt.origin), f64("upper", mt.upper)
    keep["bucket_base"] = np.ascontiguousarray(mt.bucket_base, dtype=np.int32)
    c.bucket_base = keep["bucket_base"].ctypes.data_as(c_int32_p)
    c.beta_grid, c.xi_tab = f64("beta_grid", mt.beta_grid), f64("xi_tab", mt.xi_tab)
    c.v0, c.d0, c.sv = f64("v0", mt.v0), f64("d0", mt.d0), f64("sv", mt.sv)
    c.vd_beta_dependent, c.growth_mode, c.bias = int(mt.vd_beta_dependent), int(mt.growth_mode), float(mt.bias)
    c.linear_bias = int(mt.linear_bias)
    c.template_fsigma8, c.growth_scale = float(mt.template_fsigma8), float(mt.growth_scale)
    if mt.v0b is not None:
        c.v0b, c.d0b = f64("v0b", mt.v0b), f64("d0b", mt.d0b)
    if mt.sv2d is not None:
        c.sv_ny = int(mt.sv2d.shape[1])
        c.sv2d, c.sv_ybreaks = f64("sv2d", mt.sv2d), f64("sv_ybreaks", mt.sv_ybreaks)
    else:
        c.sv_ny = 0
    c.x, c.wx = f64("x", mt.x), f64("wx", mt.wx)
    c.mu_resc, c.w_resc = f64("mu_resc", mt.mu_resc), f64("w_resc", mt.w_resc)
    return c, keep


def pack_fit(ft, s, mu, wmu):
    keep = {}

    def f64(name, a):
        keep[name] = np.ascontiguousarray(a, dtype=np.float64)
        return _dp(keep[name])

    c = FitTablesC()
    c.ns, c.npoles, c.nmu = len(s), int(wmu.shape[0]), len(mu)
    c.data_beta_dependent, c.nbeta_ccf = int(ft.data_beta_dependent), len(ft.beta_ccf)
    c.cov_fixed, c.nbeta_cov = int(ft.cov_fixed), len(ft.beta_cov)
    c.like_kind, c.use_logdet = int(ft.like_kind), int(ft.use_logdet)
    c.like_a, c.like_nm1 = float(ft.like_a), float(ft.like_nm1)
    c.s, c.mu, c.wmu = f64("s", s), f64("mu", mu), f64("wmu", wmu)
    c.beta_ccf, c.data_tab = f64("beta_ccf", ft.beta_ccf), f64("data_tab", ft.data_tab)
    c.beta_cov, c.icov = f64("beta_cov", ft.beta_cov), f64("icov", ft.icov)
    c.logdet, c.lam = f64("logdet", ft.logdet), f64("lam", ft.lam)
    return c, keep
