"""One GPU context: host tables uploaded once, then batched evaluations through the C ABI.

This is the only place the Python classes touch the device.  Inputs and outputs of the
convenience methods are numpy arrays on the host (the library stages them; copies are part of
the call); ``likelihood_ptr`` / ``theory`` with device pointers give the asynchronous,
device-resident path used by bench.py and by callers that keep their batches in torch tensors.
"""
import ctypes

import numpy as np

from . import _lib


def default_device():
    try:
        import torch
        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except ImportError:
        pass
    return 0


class Engine:
    def __init__(self, model_tables, fit=None, device=None):
        """``fit``: None or dict(ft=FitTables, s=..., mu=..., wmu=...)."""
        self.lib = _lib.load()
        if self.lib.vb200_device_count() < 1:
            raise RuntimeError("victor_b200 needs a CUDA device (B200, sm_100a); none is visible and there is "
                               f"no CPU fallback: {_lib.last_error()}")
        self.device = default_device() if device is None else int(device)
        mc, keep_m = _lib.pack_model(model_tables)
        fc = None
        self.p = None
        if fit is not None:
            fc, keep_f = _lib.pack_fit(fit["ft"], fit["s"], fit["mu"], fit["wmu"])
            self.p = int(fit["ft"].p)
        handle = ctypes.c_void_p()
        rc = self.lib.vb200_create(ctypes.byref(mc), ctypes.byref(fc) if fc is not None else None,
                                   self.device, ctypes.byref(handle))
        if rc != 0:
            msg = _lib.last_error()
            if rc == -4:
                raise NotImplementedError(msg)
            raise RuntimeError(f"vb200_create failed ({rc}): {msg}")
        self.handle = handle
        self.model_tables = model_tables

    def _check(self, rc):
        if rc != 0:
            msg = _lib.last_error()
            if rc == -4:
                raise NotImplementedError(msg)
            if rc == -1:
                raise ValueError(msg)
            raise RuntimeError(f"victor_b200 call failed ({rc}): {msg}")

    def set_option(self, key, value):
        self._check(self.lib.vb200_set_option(self.handle, key.encode(), int(value)))

    def launch_count(self):
        return int(self.lib.vb200_launch_count(self.handle))

    def synchronize(self):
        self._check(self.lib.vb200_synchronize(self.handle))

    # ---- host-array convenience paths -------------------------------------------------
    def theory(self, rows, s, mu, wmu):
        """(xi[n][nmu][ns] or None, multipoles[n][L][ns] or None); xi only when wmu is None."""
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        s = np.ascontiguousarray(s, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        n = rows.shape[0]
        xi = mult = None
        if wmu is None:
            xi = np.empty((n, len(mu), len(s)))
            L, wp = 0, None
        else:
            wmu = np.ascontiguousarray(wmu, dtype=np.float64)
            L = wmu.shape[0]
            mult = np.empty((n, L, len(s)))
            wp = wmu.ctypes.data
        self._check(self.lib.vb200_theory(
            self.handle, rows.ctypes.data, n, s.ctypes.data, len(s), mu.ctypes.data, len(mu), wp, L,
            xi.ctypes.data if xi is not None else None,
            mult.ctypes.data if mult is not None else None, None))
        return xi, mult

    def theory_pairs(self, rows, s, mu):
        """xi[n][npairs] at the separate points (s[j], mu[j])."""
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        s = np.ascontiguousarray(s, dtype=np.float64).ravel()
        mu = np.ascontiguousarray(mu, dtype=np.float64).ravel()
        if s.shape != mu.shape:
            raise ValueError("theory_pairs: s and mu must have the same number of entries")
        xi = np.empty((rows.shape[0], len(s)))
        self._check(self.lib.vb200_theory_pairs(self.handle, rows.ctypes.data, rows.shape[0], s.ctypes.data,
                                                mu.ctypes.data, len(s), xi.ctypes.data, None))
        return xi

    def likelihood(self, rows, want_theory=False):
        """(theory[n][p] or None, chi2[n], lnl[n]) as host arrays."""
        if type(rows) is not np.ndarray or rows.dtype != np.float64 or not rows.flags.c_contiguous:
            rows = np.ascontiguousarray(rows, dtype=np.float64)
        n = rows.shape[0]
        theory = np.empty((n, self.p)) if want_theory else None
        out = np.empty((2, n))          # chi2 | lnL: one allocation, two views
        base = out.ctypes.data
        rc = self.lib.vb200_likelihood(self.handle, rows.ctypes.data, n, theory.ctypes.data if want_theory else None,
                                       base, base + 8 * n, None)
        if rc != 0:
            self._check(rc)
        return theory, out[0], out[1]

    def likelihood_point(self, row):
        """(chi2, lnl) as Python floats for ONE parameter row given as a sequence of NPAR numbers: the MCMC step.
        Row and results go through two small ctypes arrays owned by the engine (no numpy allocation per call)."""
        buf = getattr(self, "_point", None)
        if buf is None:
            row_c, out_c = (ctypes.c_double * len(row))(), (ctypes.c_double * 2)()
            buf = self._point = (row_c, out_c, ctypes.addressof(row_c), ctypes.addressof(out_c))
        row_c, out_c, row_p, out_p = buf
        row_c[:] = row
        rc = self.lib.vb200_likelihood(self.handle, row_p, 1, None, out_p, out_p + 8, None)
        if rc != 0:
            self._check(rc)
        return out_c[0], out_c[1]

    # ---- raw-pointer path (host or device pointers, caller-owned buffers) ---------------
    def likelihood_ptr(self, params_ptr, n, theory_ptr, chi2_ptr, lnl_ptr, stream=None):
        self._check(self.lib.vb200_likelihood(self.handle, params_ptr, int(n), theory_ptr, chi2_ptr, lnl_ptr,
                                              stream))

    def theory_ptr(self, params_ptr, n, s, mu, wmu, xi_ptr, mult_ptr, stream=None):
        """vb200_theory with caller-owned parameter / output buffers (host or device pointers);
        ``s``, ``mu``, ``wmu`` are host arrays (staged by the library on every call)."""
        s = np.ascontiguousarray(s, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        L, wp = 0, None
        if wmu is not None:
            wmu = np.ascontiguousarray(wmu, dtype=np.float64)
            L, wp = wmu.shape[0], wmu.ctypes.data
        self._check(self.lib.vb200_theory(self.handle, params_ptr, int(n), s.ctypes.data, len(s), mu.ctypes.data,
                                          len(mu), wp, L, xi_ptr, mult_ptr, stream))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.vb200_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
