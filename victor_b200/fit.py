"""Drop-in ``CCFFit``: data vector, covariance, chi-square and log-likelihood on the B200 path.

Mirrors ``victor.CCFFit`` (victor/ccf_fit.py:10-584): same constructor dicts, attribute names
(``s, poles_s, redshift_multipoles, beta_ccf, covmat, icov, beta_covmat, fixed_data,
fixed_covmat, fit_options``), same ``chi_squared`` / ``log_likelihood`` signatures and return
values, including the reference's conventions that are part of the contract:

* the covariance / precision "interpolation" brackets with the LAST grid index
  (ccf_fit.py:225-227, 257-259), returns the end matrices outside the grid and the exact
  matrix on a grid node;
* the data vector is a PCHIP over the beta grid that extrapolates outside it (:193);
* numerical failure (NaN) is reported as ``(-inf, +inf)`` (:477-481).

``log_likelihood_batch`` evaluates a whole parameter table in one launch.
"""
import os

import numpy as np

from . import tables as _tables
from .model import CCFModel, params_to_rows, point_to_row
from .utils import InputError, load_input_file, log


class _WithLikelihood:
    """Read-only view of a model option dict with the fit's 'likelihood' entry on top (what _engine_key reads)."""
    __slots__ = ("base", "like")

    def __init__(self, base, like):
        self.base, self.like = base, like

    def __getitem__(self, key):
        return self.like if key == "likelihood" else self.base[key]

    def get(self, key, default=None):
        return self.like if key == "likelihood" else self.base.get(key, default)


class CCFFit(CCFModel):
    """Fits of the CCF model to measured multipoles (drop-in for victor.CCFFit)."""

    def __init__(self, model, data, device=None):
        # reference: ccf_fit.py:15-42
        super().__init__(model, device=device)
        base_dir = data.get("dir", "")
        data_fn = os.path.join(base_dir, data["redshift_space_ccf"].get("data_file"))
        cov_fn = os.path.join(base_dir, data["covariance_matrix"].get("data_file"))
        for fn in (data_fn, cov_fn):
            if not os.path.isfile(fn):
                raise InputError(f"Data file {fn} not found")
        self._load_redshiftspace_ccf(data["redshift_space_ccf"], data_fn)
        self._load_covariance_matrix(data["covariance_matrix"], cov_fn)
        self.fit_options = {"beta_interpolation": data.get("beta_interpolation", "datavector"),
                            "likelihood": data.get("likelihood", {"form": "Gaussian"})}

    # ------------------------------------------------------------------ loaders (host, once)
    def _load_redshiftspace_ccf(self, ccf, input_fn):
        # reference: ccf_fit.py:44-114
        input_data = load_input_file(input_fn)
        isim = ccf.get("simulation_number", None)
        if isim is not None and not isinstance(isim, int):
            raise InputError("If provided, simulation_number must be an integer")
        self.fixed_data = not ccf.get("reconstruction", False)
        if not self.fixed_data:
            beta_key = ccf.get("beta_key", None)
            if beta_key and beta_key in input_data:
                self.beta_ccf = input_data[beta_key]
                if not np.all(np.diff(self.beta_ccf) > 0):
                    raise InputError("Redshift-space beta grid must be strictly monotonically increasing")
            elif self.fixed_real_input:
                raise InputError("Reconstruction beta information required for redshift-space ccf but not found")
            else:
                self.beta_ccf = self.beta
        fmt = ccf.get("format", "multipoles")
        ccf_keys = np.atleast_1d(ccf["ccf_keys"])
        if (fmt == "multipoles" and len(ccf_keys) < 2) or (fmt == "rmu" and len(ccf_keys) != 3):
            raise InputError(f"Wrong number of redshift-space ccf keys provided for format {fmt}")
        for key in ccf_keys:
            if key not in input_data:
                raise InputError(f"Key {key} not found in file {input_fn}")
        if fmt != "multipoles":
            raise InputError("Currently only multipole format is supported for redshift-space ccf data and covmat")
        self.s = input_data[ccf_keys[0]]
        npole = len(ccf_keys) - 1
        if npole > _tables.MAX_POLES:
            raise InputError("at most three redshift-space multipoles (0, 2, 4) are supported")
        self.poles_s = np.atleast_1d([0, 2, 4][:npole])
        self.redshift_multipoles = {}
        expected = self.s.shape if self.fixed_data else (len(self.beta_ccf), len(self.s))
        for i, ell in enumerate(self.poles_s):
            arr = input_data[ccf_keys[i + 1]]
            arr = arr if isim is None else arr[isim]
            if arr.shape != expected:
                raise InputError(f"Shape of redshift ccf multipole {ell} is {arr.shape}, expected {expected}")
            self.redshift_multipoles[f"{ell}"] = arr

    def _load_covariance_matrix(self, covariance, input_fn):
        # reference: ccf_fit.py:116-164
        input_data = load_input_file(input_fn)
        if not self.fixed_data:
            self.fixed_covmat = covariance.get("fixed_beta", True)
            if not self.fixed_covmat:
                beta_key = covariance.get("beta_key", None)
                if beta_key and beta_key in input_data:
                    self.beta_covmat = input_data[beta_key]
                    if not np.all(np.diff(self.beta_covmat) > 0):
                        raise InputError("Covariance beta grid must be strictly monotonically increasing")
                else:
                    self.beta_covmat = self.beta_ccf
        else:
            self.fixed_covmat = True
        cov_key = covariance["cov_key"]
        if cov_key not in input_data:
            raise InputError(f"Key {cov_key} not found in file {input_fn}")
        covmat = input_data[cov_key]
        p = len(self.s) * len(self.poles_s)
        if self.fixed_covmat:
            if covmat.shape != (p, p):
                raise InputError("Unexpected shape of (fixed) covariance matrix")
        elif covmat.shape != (len(self.beta_covmat), p, p):
            raise InputError("Unexpected shape of (beta-varying) covariance matrix")
        self.covmat = covmat
        self.icov = np.linalg.inv(self.covmat)

    # ------------------------------------------------------------------ host-side accessors
    def get_interpolated_redshift_multipoles(self, beta=None):
        """reference: ccf_fit.py:166-193"""
        from scipy.interpolate import PchipInterpolator
        stack = np.array([self.redshift_multipoles[f"{ell}"] for ell in self.poles_s])
        if self.fixed_data:
            return np.atleast_2d(stack)
        if beta is None:
            raise InputError("Need to supply a valid value of beta for interpolation")
        return np.atleast_2d(PchipInterpolator(self.beta_ccf, stack, axis=1)(beta))

    def _bracket(self, beta):
        """(lo, hi, t) of the reference's matrix blend, ccf_fit.py:218-227 (hi = LAST index)."""
        grid = self.beta_covmat
        if beta < grid.min():
            return 0, 0, 0.0
        if beta > grid.max():
            return len(grid) - 1, len(grid) - 1, 0.0
        if beta in grid:
            i = int(np.where(grid == beta)[0][0])
            return i, i, 0.0
        lo = int(np.where(grid < beta)[0][-1])
        hi = int(np.where(grid >= beta)[0][-1])
        return lo, hi, (beta - grid[lo]) / (grid[hi] - grid[lo])

    def _blend(self, mats, beta):
        if self.fixed_covmat:
            return mats
        if beta is None:
            raise InputError("Need to supply a valid value of beta for interpolation")
        lo, hi, t = self._bracket(beta)
        if lo == hi:
            return mats[lo]
        return (1 - t) * mats[lo] + t * mats[hi]

    def get_interpolated_covariance(self, beta=None):
        """reference: ccf_fit.py:195-228"""
        return self._blend(self.covmat, beta)

    def get_interpolated_precision(self, beta=None):
        """reference: ccf_fit.py:230-260"""
        return self._blend(self.icov, beta)

    def correlation_matrix(self, beta=None):
        """reference: ccf_fit.py:262-284"""
        cov = self.get_interpolated_covariance(beta)
        d = np.sqrt(np.diag(cov))
        denom = np.outer(d, d)
        out = np.zeros_like(cov)
        np.divide(cov, denom, out=out, where=denom != 0)
        return out

    def diagonal_errors(self, beta=None):
        """reference: ccf_fit.py:286-304"""
        cov = self.get_interpolated_covariance(beta)
        return np.sqrt(np.diag(cov)).reshape((len(self.poles_s), len(self.s)))

    def multipole_datavector(self, beta=None):
        """reference: ccf_fit.py:306-323"""
        return self.get_interpolated_redshift_multipoles(beta).reshape(len(self.poles_s) * len(self.s))

    # ------------------------------------------------------------------ GPU likelihood
    def _fit_key(self, opts):
        like = opts.get("likelihood", self.fit_options["likelihood"])
        return (str(like.get("form")).lower(), like.get("nmocks", 1), like.get("nparams", None))

    def _fit_tables(self, opts):
        like = opts.get("likelihood", self.fit_options["likelihood"])
        mu, wmu = _tables.mu_projection_weights(self.poles_s, nmu=int(opts.get("mu_nodes", 100)))
        return {"ft": _tables.build_fit_tables(self, like), "s": np.asarray(self.s, dtype=np.float64),
                "mu": mu, "wmu": wmu}

    def _fit_engine(self, kwargs):
        # the reference rebuilds its options from self.model / self.fit_options on every call and feeds the
        # same kwargs to the fit options and to the model options (ccf_fit.py:379-381, 444; ccf_model.py:565-567),
        # so a user may change either dict between calls: the engine is looked up by the resolved option key
        # every time (a tuple compare; the tables are only rebuilt for a key not seen before)
        if not kwargs:   # the MCMC step: resolve the key straight from the two dicts, no copies
            key = self._engine_key(_WithLikelihood(self.model, self.fit_options["likelihood"]), True)
            eng = self._engines.get(key)
            if eng is not None:
                return eng, self.fit_options
        fit_options = dict(self.fit_options)
        fit_options.update(kwargs)
        opts = self._merged_options(kwargs)
        opts["likelihood"] = fit_options["likelihood"]
        return self._engine(opts, need_fit=True), fit_options

    def log_likelihood_batch(self, params, return_theory=False, **kwargs):
        """(lnlike[n], chisq[n]) for every parameter row, optionally with the theory vectors."""
        eng, fit_options = self._fit_engine(kwargs)
        rows = params_to_rows(params)
        if fit_options["beta_interpolation"] == "likelihood" and not self.fixed_data:
            return self._likelihood_interpolated(eng, rows, return_theory)
        theory, chi2, lnl = eng.likelihood(rows, want_theory=return_theory)
        if log.isEnabledFor(10) and len(lnl) > 1:    # logging.DEBUG
            log.debug("log_likelihood_batch: %d rows, %d failed (-inf)", len(lnl), int(np.count_nonzero(lnl == -np.inf)))
        return (lnl, chi2, theory) if return_theory else (lnl, chi2)

    def log_likelihood_device(self, params, out=None, stream=None, **kwargs):
        """Device-resident form of ``log_likelihood_batch``: returns ``(lnlike, chisq)`` as float64 CUDA
        tensors on this fit's GPU, asynchronously on ``stream`` (a ``torch.cuda.Stream``, default: the current
        one).  ``params`` is a float64 CUDA tensor ``[n, 10]`` in the row layout of ``params_to_rows`` (used in
        place) or anything ``params_to_rows`` accepts (copied to the device once).  ``out`` = a float64 CUDA
        tensor ``[2, m >= n]`` to write (lnlike | chisq) into, e.g. a slot of a gather buffer.
        Nothing is copied back to the host; torch tensors are only the buffers (data_ptr) handed to the C ABI."""
        import torch
        eng, fit_options = self._fit_engine(kwargs)
        if fit_options["beta_interpolation"] == "likelihood" and not self.fixed_data:
            raise NotImplementedError("log_likelihood_device: beta_interpolation 'likelihood' blends two evaluations "
                                      "on the host; use log_likelihood_batch")
        dev = torch.device("cuda", eng.device)
        if isinstance(params, torch.Tensor):
            rows = params
            if rows.device != dev or rows.dtype != torch.float64 or rows.dim() != 2 or rows.shape[1] != 10 \
                    or not rows.is_contiguous():
                raise ValueError(f"params tensor must be contiguous float64 [n, 10] on {dev}")
        else:
            rows = torch.from_numpy(params_to_rows(params)).to(dev, non_blocking=True)
        n = rows.shape[0]
        if out is None:
            out = torch.empty((2, n), dtype=torch.float64, device=dev)
        elif out.device != dev or out.dtype != torch.float64 or out.dim() != 2 or out.shape[0] != 2 \
                or out.shape[1] < n or out.stride(1) != 1:
            raise ValueError(f"out must be a float64 [2, m >= {n}] tensor on {dev} with contiguous rows")
        st = torch.cuda.current_stream(dev) if stream is None else stream
        if n:
            eng.likelihood_ptr(rows.data_ptr(), n, None, out[1].data_ptr(), out[0].data_ptr(), st.cuda_stream)
        return out[0, :n], out[1, :n]

    def _likelihood_interpolated(self, eng, rows, return_theory):
        """'likelihood' beta mode (ccf_fit.py:383-440): evaluate at the two bracketing grid
        values of beta and blend lnL and chi2 linearly."""
        beta = rows[:, 1]
        grid = np.asarray(self.beta_ccf)
        if np.any(~(beta > grid[0])) or np.any(beta > grid[-1]):
            raise IndexError("beta outside the redshift-space beta grid in 'likelihood' interpolation mode "
                             "(the reference raises IndexError here too)")
        lo = np.searchsorted(grid, beta, side="left") - 1
        hi = lo + 1
        t = (beta - grid[lo]) / (grid[hi] - grid[lo])
        two = np.repeat(rows, 2, axis=0)
        two[0::2, 1] = grid[lo]
        two[1::2, 1] = grid[hi]
        theory, chi2, lnl = eng.likelihood(two, want_theory=return_theory)
        lnl_b = (1 - t) * lnl[0::2] + t * lnl[1::2]
        chi2_b = (1 - t) * chi2[0::2] + t * chi2[1::2]
        # a failure at either end fails the point (ccf_fit.py:403-404)
        bad = ~np.isfinite(lnl[0::2]) | ~np.isfinite(lnl[1::2]) | np.isnan(lnl_b)
        lnl_b = np.where(bad, -np.inf, lnl_b)
        chi2_b = np.where(bad, np.inf, chi2_b)
        if return_theory:
            return lnl_b, chi2_b, theory
        return lnl_b, chi2_b

    def chi_squared(self, params, **kwargs):
        """(chisq, covmat) at one parameter point (reference: ccf_fit.py:325-354)."""
        self._check_point(params, kwargs)
        eng, _ = self._fit_engine(kwargs)
        _, chi2, _ = eng.likelihood(params_to_rows(params), want_theory=False)
        return float(chi2[0]), self.get_interpolated_covariance(params.get("beta", None))

    def log_likelihood(self, params, **kwargs):
        """(lnlike, chisq) at one parameter point (reference: ccf_fit.py:356-483)."""
        self._check_point(params, kwargs)
        row = point_to_row(params) if isinstance(params, dict) else None
        eng, fit_options = self._fit_engine(kwargs)
        if row is None or (fit_options["beta_interpolation"] == "likelihood" and not self.fixed_data):
            lnl, chi2 = self.log_likelihood_batch(params, **kwargs)
            lnl, chi2 = float(lnl[0]), float(chi2[0])
        else:
            chi2, lnl = eng.likelihood_point(row)
        if lnl == -np.inf:   # the reference prints here (ccf_fit.py:478-479)
            log.warning("Likelihood evaluation failed, returning (-inf, inf). Parameters at fail point: %s", params)
        return lnl, chi2
