"""TEST INFRASTRUCTURE ONLY - ctypes front end of oracle/table_walk.c (the plain C table walk).

``build()`` compiles the C file with gcc into ``oracle/_build/libtable_walk.so`` (git-ignored; it travels to
the GPU box with the snapshot like every other built library).  ``TableWalk(fit)`` packs the product's host
tables into the C-ABI structs of include/victor_b200.h and evaluates parameter rows on the CPU cores with
OpenMP.  Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` leg of bench.py use it.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "table_walk.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libtable_walk.so")


def build(force=False):
    """gcc -O2 -fopenmp; rebuilt when the source or the C-ABI header is newer than the library."""
    header = os.path.join(os.path.dirname(HERE), "include", "victor_b200.h")
    fresh = os.path.isfile(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(p) for p in (SRC, header))
    if fresh and not force:
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    # the system gcc first (the image's CC may point at a toolchain without libgomp); OpenMP if the compiler has it
    last = None
    for cc in ("gcc", os.environ.get("CC") or "cc"):
        for omp in (["-fopenmp"], []):
            last = subprocess.run([cc, "-O2", *omp, "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], capture_output=True, text=True)
            if last.returncode == 0:
                return LIB
    raise RuntimeError(f"could not compile {SRC}: {last.stderr[-500:]}")


class TableWalk:
    """CPU evaluation of the packed tables of a loaded ``victor_b200.CCFFit`` (or ``CCFModel``)."""

    def __init__(self, fit, options=None, likelihood=None, threads=None):
        from victor_b200 import _lib, tables as T
        path = LIB if os.path.isfile(LIB) else build()
        self.lib = ctypes.CDLL(path)
        if self.lib.tw_abi_check(ctypes.sizeof(_lib.ModelTablesC), ctypes.sizeof(_lib.FitTablesC)) != 0:
            raise ImportError("libtable_walk.so does not match include/victor_b200.h; rebuild it")
        opts = fit._merged_options(options or {})
        self.mt = T.build_model_tables(fit, opts, nx=int(opts.get("velocity_nodes", 50)))
        self.mc, self._keep_m = _lib.pack_model(self.mt)
        self.fc = None
        if hasattr(fit, "fit_options"):
            like = likelihood or fit.fit_options["likelihood"]
            ft = T.build_fit_tables(fit, like)
            mu, W = T.mu_projection_weights(fit.poles_s, nmu=int(opts.get("mu_nodes", 100)))
            self.fc, self._keep_f = _lib.pack_fit(ft, np.asarray(fit.s, float), mu, W)
            self.p = ft.p
        if threads:
            os.environ["OMP_NUM_THREADS"] = str(int(threads))
        vp = ctypes.c_void_p
        self.lib.tw_likelihood.restype = ctypes.c_int
        self.lib.tw_likelihood.argtypes = [vp, vp, vp, ctypes.c_int64, vp, vp, vp]
        self.lib.tw_theory.restype = ctypes.c_int
        self.lib.tw_theory.argtypes = [vp, vp, ctypes.c_int64, vp, ctypes.c_int32, vp, ctypes.c_int32, vp, ctypes.c_int32,
                                       vp, vp]

    @staticmethod
    def _check(rc):
        if rc == -4:
            raise NotImplementedError("table_walk.c: unsupported table set")
        if rc != 0:
            raise RuntimeError(f"table walk failed ({rc})")

    def likelihood(self, rows, want_theory=False):
        """(theory[n][p] or None, chi2[n], lnl[n]) for float64[n, NPAR] rows."""
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        n = rows.shape[0]
        theory = np.empty((n, self.p)) if want_theory else None
        chi2, lnl = np.empty(n), np.empty(n)
        self._check(self.lib.tw_likelihood(ctypes.addressof(self.mc), ctypes.addressof(self.fc), rows.ctypes.data, n,
                                           theory.ctypes.data if want_theory else None, chi2.ctypes.data, lnl.ctypes.data))
        return theory, chi2, lnl

    def theory(self, rows, s, mu, wmu=None):
        """(xi[n][nmu][ns], multipoles[n][L][ns] or None)."""
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        s = np.ascontiguousarray(s, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        n = rows.shape[0]
        xi = np.empty((n, len(mu), len(s)))
        mult, L, wp = None, 0, None
        if wmu is not None:
            wmu = np.ascontiguousarray(wmu, dtype=np.float64)
            L, wp = wmu.shape[0], wmu.ctypes.data
            mult = np.empty((n, L, len(s)))
        self._check(self.lib.tw_theory(ctypes.addressof(self.mc), rows.ctypes.data, n, s.ctypes.data, len(s),
                                       mu.ctypes.data, len(mu), wp, L, xi.ctypes.data,
                                       mult.ctypes.data if mult is not None else None))
        return xi, mult
