"""TEST INFRASTRUCTURE ONLY - CPU oracle for the victor likelihood hot path.

A numpy/scipy restatement of the reference algorithm (seshnadathur/victor 0.1.4), one
parameter point at a time, calling the same scipy routines the reference calls (FITPACK
splines, PCHIP, Simpson, norm.pdf, trapezoid).  It exists to CHECK the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  The product (``victor_b200``) never does and has no
CPU fallback.

Parity pinning: the reference ships no tests.  This oracle is pinned against outputs of the
UNMODIFIED reference run in the dev container (``oracle/make_golden.py`` ->
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` asserts agreement to ~1e-13), and the
reference itself reproduces the five chi2 / lnL pairs printed in
``notebooks/victor_usage_demo.ipynb`` cell 22.

Every function cites the reference lines (``victor/...py:a-b``) it follows.
Two scipy APIs the reference uses no longer exist in scipy >= 1.14; they are restated as:
``simps`` -> ``scipy.integrate.simpson``;  ``interp2d(x, y, z, kind)`` on a regular grid ->
``RectBivariateSpline(x, y, z.T, kx=k, ky=k, s=0)`` (the FITPACK regrid fit interp2d used).
"""
import os

import numpy as np
from scipy.integrate import quad, simpson
from scipy.interpolate import (InterpolatedUnivariateSpline, PchipInterpolator,
                               RectBivariateSpline)
from scipy.signal import savgol_filter
from scipy.special import legendre
from scipy.stats import norm

_trapz = getattr(np, "trapezoid", None) or np.trapz


class OracleInputError(Exception):
    """Mirrors victor.utils.InputError (victor/utils.py:5)."""


def _ius(x, y):
    # the reference's `_spline(..., ext=3)` everywhere on the path (ccf_model.py:17)
    return InterpolatedUnivariateSpline(x, y, ext=3)


def load_arrays(path):
    """File -> {key: array}.  Reference: ccf_model.py:54-68 (npy / hdf5); .npz added for fixtures."""
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    if path.endswith(".npy"):
        return np.load(path, allow_pickle=True).item()
    from victor_b200.io_hdf5 import read_hdf5  # file decoding only, no arithmetic
    return read_hdf5(path)


def hubble_ratio(cosmology, z):
    """E(z) = H(z)/H0 of LambdaCDM without radiation.  cosmology.py:26-45 (astropy LambdaCDM)."""
    om = cosmology.get("Omega_m", 0.31)
    ok = cosmology.get("Omega_K", 0)
    ol = 1 - om - ok
    ok0 = 1.0 - om - ol  # astropy derives curvature from Om0 + Ode0
    zp1 = 1.0 + z
    return np.sqrt(om * zp1 ** 3 + ok0 * zp1 ** 2 + ol)


def grid_interp(x, y, z, k):
    """interp2d(x, y, z, kind) replacement; returns f(xq, yq) -> array (len(yq), len(xq))."""
    x = np.asarray(x, float)
    y = np.asarray(y, float)
    spl = RectBivariateSpline(x, y, np.asarray(z, float).T, kx=k, ky=k, s=0)

    def f(xq, yq):
        xq = np.clip(np.sort(np.atleast_1d(xq)), x.min(), x.max())
        yq = np.clip(np.sort(np.atleast_1d(yq)), y.min(), y.max())
        return spl(xq, yq).T

    return f


def legendre_moments(f_rmu, r, ells, even=True, npts=200):
    """Trapezoid Legendre projection of f(r, mu).  utils.py:9-58."""
    ells = np.atleast_1d(ells)
    if even:
        mu = np.linspace(0.0, 1.0, npts)
        pref = [2 * l + 1 for l in ells]
    else:
        mu = np.linspace(-1, 1, npts)
        pref = [(2 * l + 1) / 2 for l in ells]
    out = {}
    for i, l in enumerate(ells):
        pl = legendre(l)(mu)
        vals = np.zeros(len(r))
        for j in range(len(r)):
            y = f_rmu(r[j], mu)[:, 0]
            vals[j] = pref[i] * _trapz(y * pl, mu)
        out[f"{l}"] = vals
    return out


class OracleModel:
    """State + evaluation of the model half of the path (reference class CCFModel)."""

    def __init__(self, model):
        # ccf_model.py:43-97
        self.z_eff = model["z_eff"]
        self.iaH = (1 + self.z_eff) / (100 * hubble_ratio(model.get("cosmology"), self.z_eff))
        fn = os.path.join(model.get("dir", ""), model["input_model_data_file"])
        if not os.path.isfile(fn):
            raise OracleInputError(f"File {fn} containing input model data not found")
        raw = load_arrays(fn)
        self._read_real_ccf(model["realspace_ccf"], raw)
        self.matter_model = model["matter_ccf"].get("model", "linear_bias")
        self.from_data = model["realspace_ccf"].get("from_data", False)
        if self.matter_model == "linear_bias" and not self.from_data:
            self.template_sigma8 = model["matter_ccf"].get("template_sigma8", None)
            if not self.template_sigma8:
                raise OracleInputError("template_sigma8 must be provided")
        if self.matter_model == "template":
            self._read_matter_template(model["matter_ccf"], raw)
        self._read_velocity_pdf(model["velocity_pdf"], raw)
        self.options = {
            "rsd_model": model.get("rsd_model", "streaming"),
            "kaiser_approximation": model.get("kaiser_approximation", False),
            "kaiser_coord_shift": model.get("kaiser_coord_shift", True),
            "assume_isotropic": model["realspace_ccf"].get("assume_isotropic", True),
            "realspace_ccf_from_data": self.from_data,
            "matter_model": self.matter_model,
            "bias": model["matter_ccf"].get("bias", 1.9),
            "mean_model": model["velocity_pdf"]["mean"].get("model", "linear"),
            "empirical_corr": model["velocity_pdf"]["mean"].get("empirical_corr", False),
            "velocity_independent_of_AP":
                model["velocity_pdf"].get("rescale_templates_independent_of_AP", True),
        }

    # ---- loaders ------------------------------------------------------------------------
    def _read_real_ccf(self, spec, raw):
        # ccf_model.py:99-181
        fmt = spec.get("format", "multipoles")
        self.fixed_real_input = not spec.get("reconstruction", False)
        keys = np.atleast_1d(spec["ccf_keys"])
        if not self.fixed_real_input:
            bkey = spec.get("beta_key", None)
            if bkey is None or bkey not in raw:
                raise OracleInputError("beta grid missing for reconstruction-dependent real-space ccf")
            self.beta = raw[bkey]
            if not np.all(np.diff(self.beta) > 0):
                raise OracleInputError("Realspace beta grid must be strictly increasing")
        if (fmt == "multipoles" and len(keys) < 2) or (fmt == "rmu" and len(keys) != 3):
            raise OracleInputError("Wrong number of ccf keys")
        for k in keys:
            if k not in raw:
                raise OracleInputError(f"Key {k} not found in input model data file")
        isim = spec.get("simulation_number", None)
        self.r = raw[keys[0]]
        if fmt == "rmu":
            # ccf_model.py:154-181: xi(r, mu) -> multipoles 0, 2, 4 through a (default: linear) interp2d
            mu = raw[keys[1]]
            ccf = raw[keys[2]] if isim is None else raw[keys[2]][isim]
            self.poles_r = np.array([0, 2, 4])
            if self.fixed_real_input:
                if ccf.shape != (len(self.r), len(mu)):
                    raise OracleInputError("Unexpected real-space ccf shape")
                self.real_multipoles = legendre_moments(grid_interp(self.r, mu, ccf.T, 1), self.r, self.poles_r)
            else:
                if ccf.shape != (len(self.beta), len(self.r), len(mu)):
                    raise OracleInputError("Unexpected real-space ccf shape")
                self.real_multipoles = {f"{ell}": np.zeros((len(self.beta), len(self.r))) for ell in self.poles_r}
                for i in range(len(self.beta)):
                    tmp = legendre_moments(grid_interp(self.r, mu, ccf[i].T, 1), self.r, self.poles_r)
                    for ell in self.poles_r:
                        self.real_multipoles[f"{ell}"][i] = tmp[f"{ell}"]
            return
        self.poles_r = np.atleast_1d([0, 2, 4][:len(keys) - 1])
        self.real_multipoles = {}
        for i, ell in enumerate(self.poles_r):
            arr = raw[keys[i + 1]]
            self.real_multipoles[f"{ell}"] = arr if isim is None else arr[isim]
            want = self.r.shape if self.fixed_real_input else (len(self.beta), len(self.r))
            if self.real_multipoles[f"{ell}"].shape != want:
                raise OracleInputError("Unexpected real-space multipole shape")

    def _read_matter_template(self, spec, raw):
        # ccf_model.py:183-220
        self.template_sigma8 = spec.get("template_sigma8", None)
        if not self.template_sigma8:
            raise OracleInputError("template_sigma8 must be provided")
        keys = np.atleast_1d(spec.get("template_keys"))
        if len(keys) != 2:
            raise OracleInputError("expected 2 matter template keys")
        for k in keys:
            if k not in raw:
                raise OracleInputError(f"Key {k} not found in input model data file")
        rd, dl = raw[keys[0]], raw[keys[1]]
        if len(rd) != len(dl):
            raise OracleInputError("matter template shape mismatch")
        grid = np.linspace(rd.min(), rd.max())  # 50 points
        if spec.get("integrated", False):
            self.integrated_delta = _ius(rd, dl)
            slope = np.gradient(self.integrated_delta(grid), grid)
            self.delta = _ius(grid, self.integrated_delta(grid) + grid * slope / 3)
        else:
            self.delta = _ius(rd, dl)
            enclosed = np.zeros_like(grid)
            for i in range(len(grid)):
                enclosed[i] = quad(lambda x: 3 * self.delta(x) * x ** 2 / grid[i] ** 3,
                                   0, grid[i], full_output=1)[0]
            self.integrated_delta = _ius(grid, enclosed)

    def _read_velocity_pdf(self, spec, raw):
        # ccf_model.py:222-297
        mean = spec["mean"]
        self.has_velocity_template = False
        if mean.get("model", "linear") == "template":     # :227-246
            self.template_fsigma8 = mean.get("template_fsigma8")
            if not self.template_fsigma8:
                raise OracleInputError("template_fsigma8 must be provided")
            self.z_sim = mean.get("z_sim", self.z_eff)
            self.template_hubble_ratio = mean.get("template_hubble_ratio", 1)
            vkeys = np.atleast_1d(mean.get("template_keys"))
            if len(vkeys) != 2:
                raise OracleInputError("need 2 velocity mean template keys")
            for k in vkeys:
                if k not in raw:
                    raise OracleInputError(f"Key {k} not found in input model data file")
            if len(raw[vkeys[0]]) != len(raw[vkeys[1]]):
                raise OracleInputError("mean velocity template shape mismatch")
            self.radial_velocity = _ius(raw[vkeys[0]], raw[vkeys[1]])
            self.has_velocity_template = True
        elif mean.get("model", "linear") != "linear":
            raise NotImplementedError("oracle: only the 'linear' and 'template' mean-velocity models")
        disp = spec.get("dispersion", {})
        kind = disp.get("model", "constant")
        if kind != "template":
            # ccf_model.py:284-292: the 'constant' branch leaves `sv` unbound and crashes
            raise OracleInputError("dispersion model must be 'template' (reference crashes otherwise)")
        keys = np.atleast_1d(disp.get("template_keys"))
        if len(keys) < 2 or len(keys) > 3:
            raise OracleInputError("need 2 or 3 dispersion template keys")
        for k in keys:
            if k not in raw:
                raise OracleInputError(f"Key {k} not found in input model data file")
        self.r_for_sv = raw[keys[0]]
        sv = raw[keys[-1]]
        if len(keys) == 2:
            self.mu_for_sv = np.linspace(0, 1)
            sv = (np.ones((len(self.mu_for_sv), len(self.r_for_sv))) * sv).T
        else:
            self.mu_for_sv = raw[keys[1]]
        if sv.shape != (len(self.r_for_sv), len(self.mu_for_sv)):
            raise OracleInputError("Dispersion template shape mismatch")
        if disp.get("filter", True):
            win = disp.get("filter_window", 3)
            order = disp.get("filter_order", 1)
            sv = np.array([savgol_filter(sv[:, i], win, order) for i in range(sv.shape[1])]).T
        if sv.shape[0] == len(self.r_for_sv):
            sv = sv.T
        # normalise by the monopole at the largest r (linear interp2d + 200-pt trapz), :295-297
        f = grid_interp(self.r_for_sv, self.mu_for_sv, sv, 1)
        mono = legendre_moments(f, self.r_for_sv, [0])
        self.sv_rmu = sv / mono["0"][-1]

    # ---- per-point pieces ---------------------------------------------------------------
    def real_multipoles_at(self, beta=None):
        # ccf_model.py:299-326
        stack = np.array([self.real_multipoles[f"{ell}"] for ell in self.poles_r])
        if self.fixed_real_input:
            return np.atleast_2d(stack)
        if beta is None:
            raise OracleInputError("Need a value of beta")
        return np.atleast_2d(PchipInterpolator(self.beta, stack, axis=1)(beta))

    def delta_profiles(self, r, params, opts):
        # ccf_model.py:328-383
        if opts["matter_model"] == "linear_bias":
            bias = params.get("bias", opts["bias"])
            xir = _ius(self.r, self.real_multipoles_at(params.get("beta", None))[0])
            enclosed = np.zeros_like(r)
            for i in range(len(r)):
                rr = np.linspace(0, r[i], 100)
                enclosed[i] = _trapz(xir(rr) * rr ** 2, rr)
            return xir(r) / bias, 3 * enclosed / (bias * r ** 3)
        if opts["matter_model"] == "template":
            return self.delta(r), self.integrated_delta(r)
        raise NotImplementedError(f"oracle: matter_model {opts['matter_model']}")

    def velocity_terms(self, r, params, opts):
        # ccf_model.py:385-492, linear and template mean models
        if "epsilon" in params:
            apar = params.get("alpha", 1) * params["epsilon"] ** (-2 / 3)
        else:
            apar = params.get("apar", 1)
        iaH_true = self.iaH * apar
        d_r, D_r = self.delta_profiles(r, params, opts)
        delta = _ius(r, d_r)
        Delta = _ius(r, D_r)
        if opts["matter_model"] == "linear_bias" and opts["realspace_ccf_from_data"]:
            growth = params["beta"] * params.get("bias", opts["bias"])
        else:
            growth = params["fsigma8"] / self.template_sigma8
        if opts["mean_model"] == "template":              # :439-443, 483-488
            if not self.has_velocity_template:
                raise OracleInputError("velocity_terms: no velocity template has been supplied")
            shift = (1 + self.z_sim) / (1 + self.z_eff)
            growth = (params["fsigma8"] / self.template_fsigma8) * self.template_hubble_ratio * shift / apar
            vr = self.radial_velocity(r) * growth
            rg = np.linspace(0.1, self.r.max(), 100)
            dvr = _ius(rg, np.gradient(self.radial_velocity(rg) * growth, rg))(r)
        elif not opts["empirical_corr"]:
            vr = -growth * r * Delta(r) / (3 * iaH_true)
            dvr = -growth * (delta(r) - 2 * Delta(r) / 3) / iaH_true
        else:
            Av = params.get("Av", 0)
            vr = -growth * r * Delta(r) * (1 + Av * delta(r)) / (3 * iaH_true)
            rg = np.linspace(0.1, self.r.max(), 100)
            vg = -growth * rg * Delta(rg) * (1 + Av * delta(rg)) / (3 * iaH_true)
            dvr = _ius(rg, np.gradient(vg, rg))(r)
        return vr, dvr

    def _options(self, kwargs):
        opts = dict(self.options)
        opts.update(kwargs)
        return opts

    def theory_xi(self, s, mu, params, **kwargs):
        """xi(s, mu) on the outer product of 1-D s and mu; returns (len(mu), len(s)).

        ccf_model.py:538-789.  (2-D meshgrid inputs are reduced to their unique values by the
        reference, :577; pass the 1-D grids here.)
        """
        opts = self._options(kwargs)
        rsd = opts["rsd_model"]
        x = np.linspace(-6, 6) if rsd in ("streaming", "dispersion") else 0
        S, Mu, X = np.meshgrid(np.atleast_1d(s), np.atleast_1d(mu), x)

        beta = 0.40 if (self.fixed_real_input and opts["matter_model"] != "linear_bias") else params["beta"]
        if "epsilon" in params:
            eps = params["epsilon"]
            apar = params.get("alpha", 1) * eps ** (-2 / 3)
            aperp = eps * apar
        else:
            aperp = params.get("aperp", 1)
            apar = params.get("apar", 1)
            eps = aperp / apar
        iaH_true = self.iaH * apar

        if opts["velocity_independent_of_AP"]:
            scale = params.get("astar", 1)
        else:
            mm = np.linspace(1e-10, 1)
            scale = _trapz(apar * np.sqrt(1 + (1 - mm ** 2) * (eps ** 2 - 1)), mm)
        r_ref = self.r
        r_scaled = r_ref * scale
        xi_r = self.real_multipoles_at(beta)
        xi_spl = {}
        for i, ell in enumerate(self.poles_r):
            xi_spl[f"{ell}"] = _ius(r_ref if opts["realspace_ccf_from_data"] else r_scaled, xi_r[i])
        vr, dvr = self.velocity_terms(np.append([0.01], r_ref), params, opts)
        knots_v = np.append([0.01 * scale], r_scaled)
        vr_f = _ius(knots_v, vr)
        dvr_f = _ius(knots_v, dvr / scale)
        sigma_v = params.get("sigma_v", 380)

        s_perp = S * np.sqrt(1 - Mu ** 2) * aperp
        s_par = S * Mu * apar
        s_true = np.sqrt(s_par ** 2 + s_perp ** 2)

        def xi_real(r, mu_r):
            # ccf_model.py:681-687 (template input: evaluated at the true-cosmology r, mu_r)
            if opts["assume_isotropic"]:
                return xi_spl["0"](r) * legendre(0)(mu_r)
            tot = np.zeros_like(r)
            for ell in self.poles_r:
                tot = tot + xi_spl[f"{ell}"](r) * legendre(ell)(mu_r)
            return tot

        if rsd in ("streaming", "dispersion"):
            v_par = X * sigma_v
            sv_spl = RectBivariateSpline(self.r_for_sv * scale, self.mu_for_sv, self.sv_rmu.T)
            if rsd == "streaming":
                r_par = s_par - v_par * iaH_true
                r = np.sqrt(s_perp ** 2 + r_par ** 2)
                mu_r = r_par / r
                sv = sigma_v * sv_spl.ev(r, mu_r)
                pdf = norm.pdf(v_par, loc=vr_f(r) * mu_r, scale=sv)
                jac = 1
            else:
                r_par = (s_par - v_par * iaH_true) / (1 + iaH_true * vr_f(s_true) / s_true)
                for _ in range(opts.get("niter", 5)):
                    r = np.sqrt(s_perp ** 2 + r_par ** 2)
                    r_par = (s_par - v_par * iaH_true) / (1 + iaH_true * vr_f(r) / r)
                r = np.sqrt(s_perp ** 2 + r_par ** 2)
                mu_r = r_par / r
                sv = sigma_v * sv_spl.ev(r, mu_r)
                pdf = norm.pdf(v_par, loc=0, scale=sv)
                jac = 1 / (1 + vr_f(r) * iaH_true / r + iaH_true * mu_r ** 2 * (dvr_f(r) - vr_f(r) / r))
            if opts["realspace_ccf_from_data"]:
                xi_rmu = self._xi_real_from_data(xi_spl, r_par, s_perp, apar, aperp, opts)
            else:
                xi_rmu = xi_real(r, mu_r)
            return simpson((1 + xi_rmu) * jac * pdf, x=v_par, axis=2) - 1

        if rsd in ("kaiser", "euclid_special"):
            M = params.get("M", 1.0)
            Q = params.get("Q", 1.0)
            if opts.get("kaiser_coord_shift", True):
                r_par = s_par / (1 + M * iaH_true * vr_f(s_true) / s_true)
                for _ in range(opts.get("niter", 5)):
                    r = np.sqrt(s_perp ** 2 + r_par ** 2)
                    r_par = s_par / (1 + M * iaH_true * vr_f(r) / r)
            else:
                r_par = s_par
            r = np.sqrt(s_perp ** 2 + r_par ** 2)
            mu_r = r_par / r
            a, b = (1, 1) if rsd == "kaiser" else (3, 2)
            J = a * M * vr_f(r) * iaH_true / r + b * M * Q * mu_r ** 2 * iaH_true * (dvr_f(r) - vr_f(r) / r)
            if opts["realspace_ccf_from_data"]:
                xi_rmu = self._xi_real_from_data(xi_spl, r_par, s_perp, apar, aperp, opts)
            else:
                xi_rmu = xi_real(r, mu_r)
            if rsd == "euclid_special":
                out = M * xi_rmu - J
            elif not opts.get("kaiser_approximation", False):
                out = (1 + M * xi_rmu) / (1 + J) - 1
            else:
                out = M * xi_rmu - J
            return out[:, :, 0]
        raise OracleInputError(f"Unrecognised choice of model {rsd}")

    def _xi_real_from_data(self, xi_spl, r_par, s_perp, apar, aperp, opts):
        # ccf_model.py:675-687
        rp = r_par / apar
        rt = s_perp / aperp
        r = np.sqrt(rp ** 2 + rt ** 2)
        mu_r = rp / r
        if opts["assume_isotropic"]:
            return xi_spl["0"](r) * legendre(0)(mu_r)
        tot = np.zeros_like(r)
        for ell in self.poles_r:
            tot = tot + xi_spl[f"{ell}"](r) * legendre(ell)(mu_r)
        return tot

    def theory_multipoles(self, s, params, poles=(0, 2), **kwargs):
        # ccf_model.py:791-827
        poles = np.atleast_1d(poles)
        even = not np.any(poles % 2)
        mu = np.linspace(0, 1, 100) if even else np.linspace(-1, 1, 100)
        s = np.asarray(s, float)
        xi = self.theory_xi(s, mu, params, **kwargs)
        f = grid_interp(s, mu, xi, 3)
        return legendre_moments(f, s, poles, even=even)

    def theory_multipole_vector(self, s, params, poles=(0, 2), **kwargs):
        # ccf_model.py:829-860
        mp = self.theory_multipoles(s, params, poles, **kwargs)
        return np.concatenate([mp[f"{ell}"] for ell in np.atleast_1d(poles)])


class OracleFit(OracleModel):
    """Data vector, covariance, chi-square and log-likelihood (reference class CCFFit)."""

    def __init__(self, model, data):
        super().__init__(model)
        base = data.get("dir", "")
        dfn = os.path.join(base, data["redshift_space_ccf"].get("data_file"))
        cfn = os.path.join(base, data["covariance_matrix"].get("data_file"))
        for fn in (dfn, cfn):
            if not os.path.isfile(fn):
                raise OracleInputError(f"Data file {fn} not found")
        self._read_data(data["redshift_space_ccf"], dfn)
        self._read_cov(data["covariance_matrix"], cfn)
        self.fit_options = {"beta_interpolation": data.get("beta_interpolation", "datavector"),
                            "likelihood": data.get("likelihood", {"form": "Gaussian"})}

    def _read_data(self, spec, fn):
        # ccf_fit.py:44-114
        raw = load_arrays(fn)
        isim = spec.get("simulation_number", None)
        self.fixed_data = not spec.get("reconstruction", False)
        if not self.fixed_data:
            bkey = spec.get("beta_key", None)
            if bkey and bkey in raw:
                self.beta_ccf = raw[bkey]
                if not np.all(np.diff(self.beta_ccf) > 0):
                    raise OracleInputError("Redshift-space beta grid must be strictly increasing")
            elif self.fixed_real_input:
                raise OracleInputError("beta information required for redshift-space ccf")
            else:
                self.beta_ccf = self.beta
        if spec.get("format", "multipoles") != "multipoles":
            raise OracleInputError("only multipole format is supported for redshift-space data")
        keys = np.atleast_1d(spec["ccf_keys"])
        if len(keys) < 2:
            raise OracleInputError("Wrong number of redshift-space ccf keys")
        for k in keys:
            if k not in raw:
                raise OracleInputError(f"Key {k} not found in file {fn}")
        self.s = raw[keys[0]]
        self.poles_s = np.atleast_1d([0, 2, 4][:len(keys) - 1])
        self.redshift_multipoles = {}
        for i, ell in enumerate(self.poles_s):
            arr = raw[keys[i + 1]]
            self.redshift_multipoles[f"{ell}"] = arr if isim is None else arr[isim]
            want = self.s.shape if self.fixed_data else (len(self.beta_ccf), len(self.s))
            if self.redshift_multipoles[f"{ell}"].shape != want:
                raise OracleInputError("Unexpected redshift-space multipole shape")

    def _read_cov(self, spec, fn):
        # ccf_fit.py:116-164
        raw = load_arrays(fn)
        if not self.fixed_data:
            self.fixed_covmat = spec.get("fixed_beta", True)
            if not self.fixed_covmat:
                bkey = spec.get("beta_key", None)
                if bkey and bkey in raw:
                    self.beta_covmat = raw[bkey]
                    if not np.all(np.diff(self.beta_covmat) > 0):
                        raise OracleInputError("Covariance beta grid must be strictly increasing")
                else:
                    self.beta_covmat = self.beta_ccf
        else:
            self.fixed_covmat = True
        ckey = spec["cov_key"]
        if ckey not in raw:
            raise OracleInputError(f"Key {ckey} not found in file {fn}")
        cov = raw[ckey]
        p = len(self.s) * len(self.poles_s)
        want = (p, p) if self.fixed_covmat else (len(self.beta_covmat), p, p)
        if cov.shape != want:
            raise OracleInputError("Unexpected shape of covariance matrix")
        self.covmat = cov
        self.icov = np.linalg.inv(cov)

    def data_vector(self, beta=None):
        # ccf_fit.py:166-193, 306-323
        stack = np.array([self.redshift_multipoles[f"{ell}"] for ell in self.poles_s])
        if not self.fixed_data:
            if beta is None:
                raise OracleInputError("Need a value of beta")
            stack = PchipInterpolator(self.beta_ccf, stack, axis=1)(beta)
        return np.atleast_2d(stack).reshape(len(self.poles_s) * len(self.s))

    def _blend(self, mats, beta):
        # ccf_fit.py:195-260.  NB `highind` is the LAST index with grid >= beta (reference quirk).
        if self.fixed_covmat:
            return mats
        if beta is None:
            raise OracleInputError("Need a value of beta")
        g = self.beta_covmat
        if beta < g.min():
            return mats[0]
        if beta > g.max():
            return mats[-1]
        if beta in g:
            return mats[np.where(g == beta)[0][0]]
        lo = np.where(g < beta)[0][-1]
        hi = np.where(g >= beta)[0][-1]
        t = (beta - g[lo]) / (g[hi] - g[lo])
        return (1 - t) * mats[lo] + t * mats[hi]

    def covariance_at(self, beta=None):
        return self._blend(self.covmat, beta)

    def precision_at(self, beta=None):
        return self._blend(self.icov, beta)

    def chi_squared(self, params, **kwargs):
        # ccf_fit.py:325-354
        th = self.theory_multipole_vector(self.s, params, self.poles_s, **kwargs)
        b = params.get("beta", None)
        resid = th - self.data_vector(b)
        return np.dot(np.dot(resid, self.precision_at(b)), resid), self.covariance_at(b)

    def _form(self, chisq, norm_term, like):
        # ccf_fit.py:455-473
        form = like["form"].lower()
        nm = like.get("nmocks", 1)
        if form == "sellentin":
            return -nm * np.log(1 + chisq / (nm - 1)) / 2 + norm_term
        if form == "hartlap":
            p = len(self.s) * len(self.poles_s)
            return -0.5 * chisq * (nm - p - 2) / (nm - 1) + norm_term
        if form == "percival":
            npar = like["nparams"]
            nd = len(self.s) * len(self.poles_s)
            B = (nm - nd - 2) / ((nm - nd - 1) * (nm - nd - 4))
            m = npar + 2 + (nm - 1 + B * (nd - npar)) / (1 + B * (nd - npar))
            return -m * np.log(1 + chisq / (nm - 1)) / 2 + norm_term
        if form == "gaussian":
            return -0.5 * chisq + norm_term
        raise OracleInputError("Unrecognised likelihood form")

    def log_likelihood(self, params, **kwargs):
        # ccf_fit.py:356-483
        fo = dict(self.fit_options)
        fo.update(kwargs)
        like = fo["likelihood"]
        if fo["beta_interpolation"] == "likelihood" and not self.fixed_data:
            beta = params["beta"]
            lo = np.where(self.beta_ccf < beta)[0][-1]
            hi = np.where(self.beta_ccf >= beta)[0][0]
            t = (beta - self.beta_ccf[lo]) / (self.beta_ccf[hi] - self.beta_ccf[lo])
            ends = []
            for idx in (lo, hi):
                p = dict(params)
                p["beta"] = self.beta_ccf[idx]
                c2, cov = self.chi_squared(p, **kwargs)
                if not self.fixed_covmat:
                    sign, ld = np.linalg.slogdet(cov)
                    if sign != 1:
                        return -np.inf, np.inf
                    nt = -0.5 * ld
                else:
                    nt = 0
                ends.append((c2, self._form(c2, nt, like)))
            lnl = (1 - t) * ends[0][1] + t * ends[1][1]
            chisq = (1 - t) * ends[0][0] + t * ends[1][0]
        else:
            chisq, cov = self.chi_squared(params, **kwargs)
            if not self.fixed_covmat:
                sign, ld = np.linalg.slogdet(cov)
                if sign != 1:
                    return -np.inf, np.inf
                nt = -0.5 * ld
            else:
                nt = 0
            lnl = self._form(chisq, nt, like)
        if np.isnan(lnl):
            return -np.inf, np.inf
        return lnl, chisq
