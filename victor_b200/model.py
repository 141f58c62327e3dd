"""Drop-in ``CCFModel``: victor's model class served by the B200 CUDA path.

Mirrors the public surface of ``victor.CCFModel`` (victor/ccf_model.py:24-1041) for the
likelihood hot path: same constructor dict, same attribute names (``r, beta, poles_r, iaH,
r_for_sv, mu_for_sv, sv_rmu, template_sigma8, model`` ...), same method names and return
types.  Construction (file loading, template preparation) runs on the host with the same
scipy routines as the reference; every evaluation runs on the GPU through the C-ABI library
(``victor_b200/csrc``) -- there is no CPU evaluation path.

New, batched entry points: ``theory_multipole_vector_batch`` / ``theory_multipoles_batch`` /
``theory_xi_batch`` take a parameter table (dict of arrays or float64[n, k]) and evaluate all
rows in one launch; the reference-style single-point methods call them with n = 1.
"""
import os

import numpy as np
from scipy.integrate import quad
from scipy.interpolate import InterpolatedUnivariateSpline, RectBivariateSpline
from scipy.special import legendre

from . import tables as _tables
from .utils import GridInterpolator2D, InputError, load_input_file, log, trapezoid


def _ext3_spline(x, y):
    return InterpolatedUnivariateSpline(x, y, ext=3)


def _grid_multipoles(r, mu, f_rmu, ells, npts=200):
    """Legendre multipoles at ``r`` of a function tabulated as f_rmu[len(r)][len(mu)], even in mu.

    The reference's ``utils.multipoles_from_fn(interp2d(r, mu, f.T), r, ell)`` (utils.py:9-58): bilinear
    interpolation (interp2d's default kind) to 200 mu values in [0, 1] -- arguments outside the table are
    moved to its edge -- then (2l+1) x trapezoid of f L_l."""
    r, mu = np.asarray(r, float), np.asarray(mu, float)
    lin = RectBivariateSpline(r, mu, np.asarray(f_rmu, float), kx=1, ky=1, s=0)
    fine = np.linspace(0.0, 1.0, npts)
    rows = lin(r, np.clip(fine, mu.min(), mu.max()))            # [len(r)][npts]
    return {f"{ell}": np.array([(2 * ell + 1) * trapezoid(row * legendre(ell)(fine), fine) for row in rows])
            for ell in np.atleast_1d(ells)}


def hubble_ratio(cosmology, z):
    """E(z) for LambdaCDM with Om, Ok, OL = 1 - Om - Ok and no radiation.

    victor/cosmology.py:26-45 builds ``astropy.cosmology.LambdaCDM(H0, Om0, Ode0)`` (whose
    default Tcmb0 = 0 means no radiation term) and returns H(z)/H0.
    """
    cosmology = cosmology or {}
    om = cosmology.get("Omega_m", 0.31)
    ok = cosmology.get("Omega_K", 0)
    ol = 1 - om - ok
    curv = 1.0 - om - ol
    zp1 = 1.0 + z
    return float(np.sqrt(om * zp1 ** 3 + curv * zp1 ** 2 + ol))


# defaults of the columns a parameter table may leave out, in tables.PARAM_ORDER (bias NaN: the model's own, :359)
_DEFAULT_ROW = np.array([np.nan, np.nan, 380.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.0, np.nan])


def point_to_row(params):
    """One parameter point given as a dict of plain numbers -> the list of its NPAR row values (the MCMC step's
    fast path), or None if some value is not a scalar.  Same conventions as ``params_to_rows``."""
    get = params.get
    try:
        fs8, beta, sig = float(get("fsigma8", np.nan)), float(get("beta", np.nan)), float(get("sigma_v", 380.0))
        if "epsilon" in params:
            eps = float(params["epsilon"])
            apar = float(get("alpha", 1.0)) * eps ** (-2 / 3)
            aperp = eps * apar
        else:
            aperp, apar = float(get("aperp", 1.0)), float(get("apar", 1.0))
        return [fs8, beta, sig, aperp, apar, float(get("astar", 1.0)), float(get("M", 1.0)), float(get("Q", 1.0)),
                float(get("Av", 0.0)), float(get("bias", np.nan))]
    except TypeError:      # some value is an array
        return None


def params_to_rows(params, n_hint=None):
    """Normalise a parameter specification to float64[n, NPAR] (tables.PARAM_ORDER).

    Accepts a dict of scalars/arrays with victor's parameter names, or an array [n, k] whose
    first columns follow PARAM_ORDER (k >= 3; missing columns take victor's defaults).  The
    (epsilon, alpha) parameterisation is converted exactly as the reference does
    (ccf_model.py:589-596): apar = alpha * epsilon^(-2/3), aperp = epsilon * apar.
    """
    defaults = {"fsigma8": np.nan, "beta": np.nan, "sigma_v": 380.0, "aperp": 1.0, "apar": 1.0,
                "astar": 1.0, "M": 1.0, "Q": 1.0, "Av": 0.0, "bias": np.nan}   # bias NaN: the model's own (:359)
    if isinstance(params, dict):
        # fast path for one parameter point given as plain numbers (the MCMC step): no array work
        row = point_to_row(params)
        if row is not None and (not n_hint or n_hint == 1):
            return np.array([row], dtype=np.float64)
        cols = {}
        lens = [np.size(v) for k, v in params.items()
                if k in defaults or k in ("epsilon", "alpha")]
        n = max(lens) if lens else 1
        if n_hint:
            n = max(n, n_hint)

        def col(name, default):
            v = params.get(name, default)
            return np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)).copy()

        if "epsilon" in params:
            eps = col("epsilon", 1.0)
            # Python-float pow for small tables: bit-identical to the reference's scalar
            # `epsilon**(-2/3)`; numpy's vectorised pow may differ from libm in the last bit
            powed = np.array([float(e) ** (-2 / 3) for e in eps]) if n <= 1024 else eps ** (-2 / 3)
            cols["apar"] = col("alpha", 1.0) * powed
            cols["aperp"] = eps * cols["apar"]
        rows = np.empty((n, _tables.NPAR))
        for i, name in enumerate(_tables.PARAM_ORDER):
            rows[:, i] = cols[name] if name in cols else col(name, defaults[name])
        return rows
    arr = np.atleast_2d(np.asarray(params, dtype=np.float64))
    if arr.shape[1] == _tables.NPAR:
        return np.ascontiguousarray(arr)
    if arr.shape[1] < 3 or arr.shape[1] > _tables.NPAR:
        raise InputError(f"parameter table must have between 3 and {_tables.NPAR} columns "
                         f"{_tables.PARAM_ORDER}")
    rows = np.empty((arr.shape[0], _tables.NPAR))
    k = arr.shape[1]
    rows[:, :k] = arr
    rows[:, k:] = _DEFAULT_ROW[k:]
    return rows


class CCFModel:
    """Redshift-space void-galaxy / density-split CCF model (drop-in for victor.CCFModel)."""

    extensions = {"npy": [".npy"], "npz": [".npz"],
                  "hdf5": [".hdf", ".h4", ".hdf4", ".he2", ".h5", ".hdf5", ".he5", ".h5py"]}

    def __init__(self, model, device=None):
        # reference: ccf_model.py:33-97
        self.z_eff = model["z_eff"]
        self.iaH = (1 + self.z_eff) / (100 * hubble_ratio(model.get("cosmology"), self.z_eff))

        input_fn = os.path.join(model.get("dir", ""), model["input_model_data_file"])
        if not os.path.isfile(input_fn):
            raise InputError(f"File {input_fn} containing input model data not found")
        input_data = load_input_file(input_fn)

        self._load_realspace_ccf(model["realspace_ccf"], input_data)
        self.template_sigma8 = None
        self.matter_model = model["matter_ccf"].get("model", "linear_bias")
        self.realspace_ccf_from_data = model["realspace_ccf"].get("from_data", False)
        if self.matter_model == "linear_bias" and not self.realspace_ccf_from_data:
            self.template_sigma8 = model["matter_ccf"].get("template_sigma8", None)
            if not self.template_sigma8:
                raise InputError("When using linear bias for the matter ccf and the real-space ccf is from "
                                 "a template, template_sigma8 must be provided")
        if self.matter_model == "template":
            self._set_matter_ccf_template(model["matter_ccf"], input_data)
        self._set_velocity_pdf(model["velocity_pdf"], input_data)

        self.model = {
            "rsd_model": model.get("rsd_model", "streaming"),
            "kaiser_approximation": model.get("kaiser_approximation", False),
            "kaiser_coord_shift": model.get("kaiser_coord_shift", True),
            "assume_isotropic": model["realspace_ccf"].get("assume_isotropic", True),
            "realspace_ccf_from_data": self.realspace_ccf_from_data,
            "matter_model": self.matter_model,
            "excursion_set_options": model["matter_ccf"].get("excursion_set_options", {}),
            "bias": model["matter_ccf"].get("bias", 1.9),
            "mean_model": model["velocity_pdf"]["mean"].get("model", "linear"),
            "pdf_form": model["velocity_pdf"].get("form", "gaussian"),
            "empirical_corr": model["velocity_pdf"]["mean"].get("empirical_corr", False),
            "velocity_independent_of_AP":
                model["velocity_pdf"].get("rescale_templates_independent_of_AP", True),
            # quadrature sizes: the reference hard-codes 50 velocity nodes (ccf_model.py:570) and 100
            # mu nodes (:819, :822); kept as options so that denser grids can be requested
            "velocity_nodes": model.get("velocity_nodes", 50),
            "mu_nodes": model.get("mu_nodes", 100),
        }
        self._device = device
        self._engines = {}

    # ------------------------------------------------------------------ loaders (host, once)
    def _load_realspace_ccf(self, realspace_ccf, input_data):
        # reference: ccf_model.py:99-181
        fmt = realspace_ccf.get("format", "multipoles")
        self.fixed_real_input = not realspace_ccf.get("reconstruction", False)
        ccf_keys = np.atleast_1d(realspace_ccf["ccf_keys"])
        if not self.fixed_real_input:
            beta_key = realspace_ccf.get("beta_key", None)
            if beta_key is None:
                raise InputError("Reconstruction specified for realspace ccf but no beta key provided")
            if beta_key not in input_data:
                raise InputError(f"Key {beta_key} not found in input model data file")
            self.beta = input_data[beta_key]
            if not np.all(np.diff(self.beta) > 0):
                raise InputError("Realspace beta grid must be strictly monotonically increasing")
        if (fmt == "multipoles" and len(ccf_keys) < 2) or (fmt == "rmu" and len(ccf_keys) != 3):
            raise InputError(f"Wrong number of ccf keys provided for ccf format {fmt}")
        for key in ccf_keys:
            if key not in input_data:
                raise InputError(f"Key {key} not found in input model data file")
        isim = realspace_ccf.get("simulation_number", None)
        if isim is not None and not isinstance(isim, int):
            raise InputError("If provided, simulation_number must be an integer")
        if fmt not in ("multipoles", "rmu"):
            raise InputError(f"Unrecognised real-space ccf format {fmt}: options are 'multipoles' or 'rmu'")
        self.r = input_data[ccf_keys[0]]
        if fmt == "rmu":
            # xi(r, mu) on a grid -> multipoles 0, 2, 4 by bilinear interpolation and a 200-point trapezoid
            # over mu in [0, 1] (ccf_model.py:154-181: interp2d(r, mu, xi.T) with its default kind='linear',
            # then utils.multipoles_from_fn, utils.py:45-56)
            mu = np.asarray(input_data[ccf_keys[1]], float)
            real_ccf = input_data[ccf_keys[2]]
            real_ccf = real_ccf if isim is None else real_ccf[isim]
            self.poles_r = np.array([0, 2, 4])
            if self.fixed_real_input:
                if real_ccf.shape != (len(self.r), len(mu)):
                    raise InputError(f"Shape of real ccf is {real_ccf.shape}, expected ({len(self.r)}, {len(mu)})")
                self.real_multipoles = _grid_multipoles(self.r, mu, real_ccf, self.poles_r)
            else:
                want = (len(self.beta), len(self.r), len(mu))
                if real_ccf.shape != want:
                    raise InputError(f"Shape of real ccf is {real_ccf.shape}, expected {want}")
                per_beta = [_grid_multipoles(self.r, mu, real_ccf[i], self.poles_r) for i in range(len(self.beta))]
                self.real_multipoles = {f"{ell}": np.array([pb[f"{ell}"] for pb in per_beta]) for ell in self.poles_r}
            return
        npole = len(ccf_keys) - 1
        self.poles_r = np.atleast_1d([0, 2, 4][:npole])
        self.real_multipoles = {}
        expected = self.r.shape if self.fixed_real_input else (len(self.beta), len(self.r))
        for i, ell in enumerate(self.poles_r):
            arr = input_data[ccf_keys[i + 1]]
            arr = arr if isim is None else arr[isim]
            if arr.shape != expected:
                raise InputError(f"Shape of real ccf multipole {ell} is {arr.shape}, expected {expected}")
            self.real_multipoles[f"{ell}"] = arr

    def _set_matter_ccf_template(self, matter_ccf, input_data):
        # reference: ccf_model.py:183-220
        self.template_sigma8 = matter_ccf.get("template_sigma8", None)
        if not self.template_sigma8:
            raise InputError("When using template model for the matter ccf, template_sigma8 must be provided")
        template_keys = np.atleast_1d(matter_ccf.get("template_keys"))
        if len(template_keys) != 2:
            raise InputError("Wrong number of matter ccf template keys provided: expected 2 "
                             "(radial distance and monopole)")
        for key in template_keys:
            if key not in input_data:
                raise InputError(f"Key {key} not found in input model data file")
        r_for_delta = input_data[template_keys[0]]
        delta = input_data[template_keys[1]]
        if len(r_for_delta) != len(delta):
            raise InputError(f"Shape of matter ccf template is {len(delta)}, expected {len(r_for_delta)}")
        grid = np.linspace(r_for_delta.min(), r_for_delta.max())
        if matter_ccf.get("integrated", False):
            self.integrated_delta = _ext3_spline(r_for_delta, delta)
            slope = np.gradient(self.integrated_delta(grid), grid)
            self.delta = _ext3_spline(grid, self.integrated_delta(grid) + grid * slope / 3)
        else:
            self.delta = _ext3_spline(r_for_delta, delta)
            enclosed = np.zeros_like(grid)
            for i, rr in enumerate(grid):
                enclosed[i] = quad(lambda x: 3 * self.delta(x) * x ** 2 / rr ** 3, 0, rr, full_output=1)[0]
            self.integrated_delta = _ext3_spline(grid, enclosed)

    def _set_velocity_pdf(self, velocity_pdf, input_data):
        # reference: ccf_model.py:222-297
        mean_model = velocity_pdf["mean"].get("model", "linear")
        self.has_velocity_template = False
        if mean_model == "template":  # a testing option of the reference (ccf_model.py:227-246)
            mean = velocity_pdf["mean"]
            self.template_fsigma8 = mean.get("template_fsigma8")
            if not self.template_fsigma8:
                raise InputError("When using template model for the mean of the velocity pdf, a value for "
                                 "template_fsigma8 must be provided")
            self.z_sim = mean.get("z_sim", self.z_eff)
            self.template_hubble_ratio = mean.get("template_hubble_ratio", 1)
            template_keys = np.atleast_1d(mean.get("template_keys"))
            if len(template_keys) != 2:
                raise InputError(f"{len(template_keys)} velocity mean template keys provided, require 2")
            for key in template_keys:
                if key not in input_data:
                    raise InputError(f"Key {key} not found in input model data file")
            r_for_v, vr = input_data[template_keys[0]], input_data[template_keys[1]]
            if len(r_for_v) != len(vr):
                raise InputError(f"Shape of mean velocity template is {len(vr)}, expected {len(r_for_v)}")
            self.radial_velocity = _ext3_spline(r_for_v, vr)
            self.has_velocity_template = True
        if mean_model == "nonlinear" and self.matter_model != "excursion_set":
            raise InputError("Cannot have nonlinear mean velocity model unless using excursion_set matter model")
        dispersion = velocity_pdf.get("dispersion", {})
        disp_model = dispersion.get("model", "constant")
        if disp_model == "template":
            template_keys = np.atleast_1d(dispersion.get("template_keys"))
            if len(template_keys) < 2 or len(template_keys) > 3:
                raise InputError(f"{len(template_keys)} velocity dispersion template keys provided, require 2 or 3")
            for key in template_keys:
                if key not in input_data:
                    raise InputError(f"Key {key} not found in input model data file")
            self.r_for_sv = input_data[template_keys[0]]
            sv = input_data[template_keys[-1]]
            self.sv_isotropic = len(template_keys) == 2
            if self.sv_isotropic:
                self.mu_for_sv = np.linspace(0, 1)
                sv = (np.ones((len(self.mu_for_sv), len(self.r_for_sv))) * sv).T
            else:
                self.mu_for_sv = input_data[template_keys[1]]
            if sv.shape != (len(self.r_for_sv), len(self.mu_for_sv)):
                raise InputError(f"Dispersion template shape {sv.shape} does not match expected "
                                 f"({len(self.r_for_sv)}, {len(self.mu_for_sv)})")
            if dispersion.get("filter", True):
                from scipy.signal import savgol_filter
                window = dispersion.get("filter_window", 3)
                polyorder = dispersion.get("filter_order", 1)
                sv = np.array([savgol_filter(sv[:, i], window, polyorder) for i in range(sv.shape[1])]).T
        elif disp_model == "constant":
            # the reference leaves `sv` unbound on this branch and raises UnboundLocalError
            # (ccf_model.py:284-292); report it as an input problem instead
            raise InputError("dispersion model 'constant' is not usable (the reference fails on it too); "
                             "provide a dispersion template")
        else:
            raise InputError(f"Bad choice '{disp_model}' for dispersion model, options are 'constant' or 'template'")
        if sv.shape[0] == len(self.r_for_sv):
            sv = sv.T
        # normalise by the monopole at the largest r: bilinear interp2d + 200-point trapezoid
        # over mu in [0, 1] (ccf_model.py:295-297, utils.py:45-56)
        lin = RectBivariateSpline(self.r_for_sv, self.mu_for_sv, sv.T, kx=1, ky=1, s=0)
        mu = np.linspace(0.0, 1.0, 200)
        row = lin(self.r_for_sv[-1], mu)[0]
        monopole_last = 1 * trapezoid(row * legendre(0)(mu), mu)
        self.sv_rmu = sv / monopole_last

    # ------------------------------------------------------------------ option handling
    def _merged_options(self, kwargs):
        opts = dict(self.model)
        opts.update(kwargs)
        return opts

    def _engine_key(self, opts, need_fit):
        return (opts["rsd_model"], bool(opts["assume_isotropic"]), bool(opts["velocity_independent_of_AP"]),
                opts["matter_model"], opts["mean_model"], bool(opts["empirical_corr"]),
                bool(opts["realspace_ccf_from_data"]), bool(opts.get("kaiser_approximation", False)),
                bool(opts.get("kaiser_coord_shift", True)), int(opts.get("velocity_nodes", 50)),
                int(opts.get("mu_nodes", 100)), float(opts.get("bias", 1.9)), int(opts.get("niter", 5)),
                self._fit_key(opts) if need_fit else None)

    def _engine(self, opts, need_fit=False):
        """Context on the GPU for this option set (tables are built and uploaded once per set)."""
        from .engine import Engine
        key = self._engine_key(opts, need_fit)
        eng = self._engines.get(key)
        if eng is None:
            mt = _tables.build_model_tables(self, opts, nx=int(opts.get("velocity_nodes", 50)))
            fit = self._fit_tables(opts) if need_fit else None
            eng = Engine(mt, fit, device=self._device)
            log.info("GPU context on device %s: rsd_model=%s, %d cells, %d real-space pole(s), beta-dependent=%s%s",
                     eng.device, opts["rsd_model"], mt.ncell, mt.n_ell, mt.beta_dependent,
                     ", with likelihood tables" if need_fit else "")
            self._engines[key] = eng
        return eng

    def _fit_key(self, opts):
        return None

    def _fit_tables(self, opts):
        return None

    # ------------------------------------------------------------------ evaluation (GPU)
    def get_interpolated_real_multipoles(self, beta=None):
        """Host helper with the reference's semantics (ccf_model.py:299-326); not on the GPU path."""
        from scipy.interpolate import PchipInterpolator
        stack = np.array([self.real_multipoles[f"{ell}"] for ell in self.poles_r])
        if self.fixed_real_input:
            return np.atleast_2d(stack)
        if beta is None:
            raise InputError("Need to supply a valid value of beta for interpolation")
        return np.atleast_2d(PchipInterpolator(self.beta, stack, axis=1)(beta))

    def delta_profiles(self, r, params, **kwargs):
        """Matter ccf monopole delta(r) and its volume average Delta(r) at ``r``: a host helper with the
        reference's semantics (ccf_model.py:328-383), for plots and checks; the GPU path has these
        folded into its velocity tables."""
        opts = self._merged_options(kwargs)
        r = np.asarray(r, dtype=np.float64)
        if opts["matter_model"] == "template":
            return self.delta(r), self.integrated_delta(r)
        if opts["matter_model"] == "linear_bias":
            bias = params.get("bias", opts["bias"])
            xi0 = _ext3_spline(self.r, self.get_interpolated_real_multipoles(params.get("beta", None))[0])
            enclosed = np.zeros_like(r)
            for i, ri in enumerate(r):                       # 100-point trapezoid of xi_0 r'^2 on [0, r]
                grid = np.linspace(0, ri, 100)
                enclosed[i] = trapezoid(xi0(grid) * grid ** 2, grid)
            return xi0(r) / bias, 3 * enclosed / (bias * r ** 3)
        if opts["matter_model"] == "excursion_set":
            raise NotImplementedError("matter_model 'excursion_set' is outside the B200 path")
        raise InputError(f"Invalid choice of matter_model {opts['matter_model']}")

    def velocity_terms(self, r, params, **kwargs):
        """Mean radial velocity v_r(r) and dv_r/dr at ``r``: a host helper with the reference's semantics
        (ccf_model.py:385-492) for the 'linear' (optionally empirically corrected) and 'template' mean
        models; the GPU path evaluates the same profiles from its tables."""
        opts = self._merged_options(kwargs)
        r = np.asarray(r, dtype=np.float64)
        apar = params.get("alpha", 1) * params["epsilon"] ** (-2 / 3) if "epsilon" in params else params.get("apar", 1)
        iaH_true = self.iaH * apar
        d_r, D_r = self.delta_profiles(r, params, **kwargs)
        delta, Delta = _ext3_spline(r, d_r), _ext3_spline(r, D_r)
        fine = np.linspace(0.1, self.r.max(), 100)

        def slope(values):
            return _ext3_spline(fine, np.gradient(values, fine))(r)

        if opts["mean_model"] == "template":
            if not self.has_velocity_template:
                raise InputError("velocity_terms: Cannot use template option as no template has been supplied.")
            growth = ((params["fsigma8"] / self.template_fsigma8) * self.template_hubble_ratio
                      * ((1 + self.z_sim) / (1 + self.z_eff)) / apar)
            return self.radial_velocity(r) * growth, slope(self.radial_velocity(fine) * growth)
        if opts["mean_model"] != "linear":
            raise NotImplementedError(f"mean-velocity model '{opts['mean_model']}' is outside the B200 path")
        if opts["matter_model"] == "linear_bias" and opts["realspace_ccf_from_data"]:
            growth = params["beta"] * params.get("bias", opts["bias"])
        else:
            growth = params["fsigma8"] / self.template_sigma8
        if not opts["empirical_corr"]:
            return (-growth * r * Delta(r) / (3 * iaH_true),
                    -growth * (delta(r) - 2 * Delta(r) / 3) / iaH_true)
        Av = params.get("Av", 0)
        vr = -growth * r * Delta(r) * (1 + Av * delta(r)) / (3 * iaH_true)
        return vr, slope(-growth * fine * Delta(fine) * (1 + Av * delta(fine)) / (3 * iaH_true))

    def theory_xi_pairs_batch(self, s, mu, params, **kwargs):
        """xi at the separate points (s[j], mu[j]) for every parameter row: float64[n, len(s)]."""
        opts = self._merged_options(kwargs)
        return self._engine(opts).theory_pairs(params_to_rows(params), s, mu)

    @staticmethod
    def _sky_grid(rmax):
        # ccf_model.py:883-888: 50 x 50 grid in (s_perp, s_par), s_par on both sides of zero
        sperp = np.linspace(0.01, rmax)
        spar = np.linspace(-rmax, rmax)
        sigma, pi = np.meshgrid(sperp, spar)
        s = np.sqrt(sigma ** 2 + pi ** 2)
        return sperp, spar, s, pi / s

    def theory_xi_2D(self, params, rmax=85, **kwargs):
        """Model xi^s(s_perp, s_par) out to ``rmax`` in each direction as an interpolating function
        ``f(s_perp, s_par)`` (reference: ccf_model.py:862-894).  The reference fills its 50 x 50 grid with
        2500 scalar ``theory_xi`` calls; here all 2500 (s, mu) points go through one kernel launch."""
        self._check_point(params, kwargs)
        sperp, spar, s, mu = self._sky_grid(rmax)
        xi = self.theory_xi_pairs_batch(s.ravel(), mu.ravel(), params, **kwargs)[0].reshape(s.shape)
        return GridInterpolator2D(sperp, spar, xi)

    def xi_2D_from_multipoles(self, params, rmax=85, **kwargs):
        """xi(s_perp, s_par) rebuilt from the model multipoles 0, 2, 4 (reference: ccf_model.py:896-934)."""
        s1d = np.linspace(0.01, rmax)
        mult = self.theory_multipoles(s1d, params, poles=[0, 2, 4], **kwargs)
        sperp, spar, s, mu = self._sky_grid(rmax)
        grid = np.zeros_like(s)
        for ell in (0, 2, 4):
            grid = grid + InterpolatedUnivariateSpline(s1d, mult[f"{ell}"])(s) * legendre(ell)(mu)
        return GridInterpolator2D(sperp, spar, grid)

    def theory_xi_batch(self, s, mu, params, **kwargs):
        """xi(s, mu) for every parameter row: float64[n, len(mu), len(s)]."""
        opts = self._merged_options(kwargs)
        s = np.atleast_1d(np.asarray(s, dtype=np.float64))
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        return self._engine(opts).theory(params_to_rows(params), s, mu, None)[0]

    def theory_xi(self, s, mu, params, **kwargs):
        """Model xi(s, mu) (reference: ccf_model.py:538-789).

        1-D ``s`` and ``mu`` give an array of shape (len(mu), len(s)); 2-D meshgrid inputs are
        reduced to their sorted unique values first, as the reference does (:577).
        """
        s = np.atleast_1d(s)
        mu = np.atleast_1d(mu)
        if np.ndim(s) == 2 and np.ndim(mu) == 2:
            if s.shape != mu.shape:
                raise InputError("theory_xi: If arguments s and mu are 2D arrays they must have same shape")
            s, mu = np.unique(s), np.unique(mu)
        elif not (np.ndim(s) == 1 and np.ndim(mu) == 1):
            raise InputError("theory_xi: arguments s and mu have incompatible dimensions")
        self._check_point(params, kwargs)
        return self.theory_xi_batch(s, mu, params, **kwargs)[0]

    def theory_multipoles_batch(self, s, params, poles=(0, 2), **kwargs):
        """Legendre multipoles for every parameter row: float64[n, len(poles), len(s)]."""
        opts = self._merged_options(kwargs)
        poles = np.atleast_1d(poles)
        s = np.atleast_1d(np.asarray(s, dtype=np.float64))
        if len(s) < 4:
            raise InputError("theory_multipoles needs at least 4 values of s (bicubic interpolation in the "
                             "reference, ccf_model.py:824)")
        if np.any(np.diff(s) <= 0):
            raise InputError("theory_multipoles: s must be strictly increasing")
        mu, W = _tables.mu_projection_weights(poles, nmu=int(opts.get("mu_nodes", 100)))
        eng, rows = self._engine(opts), params_to_rows(params)
        # the kernels project onto at most MAX_POLES multipoles per launch; longer lists go in groups
        groups = [eng.theory(rows, s, mu, W[a:a + _tables.MAX_POLES])[1] for a in range(0, len(poles), _tables.MAX_POLES)]
        return groups[0] if len(groups) == 1 else np.concatenate(groups, axis=1)

    def theory_multipoles(self, s, params, poles=[0, 2], **kwargs):
        """Dict of model multipoles at ``s`` (reference: ccf_model.py:791-827)."""
        self._check_point(params, kwargs)
        poles = np.atleast_1d(poles)
        out = self.theory_multipoles_batch(s, params, poles, **kwargs)[0]
        return {f"{ell}": out[i] for i, ell in enumerate(poles)}

    def theory_multipole_vector_batch(self, s, params, poles=(0, 2), **kwargs):
        """Stacked theory vectors, float64[n, len(poles) * len(s)]."""
        out = self.theory_multipoles_batch(s, params, poles, **kwargs)
        return out.reshape(out.shape[0], -1)

    def theory_multipole_vector(self, s, params, poles=[0, 2], **kwargs):
        """Theory vector stacking all requested multipoles (reference: ccf_model.py:829-860)."""
        self._check_point(params, kwargs)
        return self.theory_multipole_vector_batch(s, params, poles, **kwargs)[0]

    def _check_point(self, params, kwargs):
        """Reference behaviour for missing parameters: KeyError on params['beta'] / ['fsigma8']."""
        if isinstance(params, dict):
            matter = kwargs["matter_model"] if "matter_model" in kwargs else self.model["matter_model"]
            if not (self.fixed_real_input and matter != "linear_bias"):
                params["beta"]
            params["fsigma8"]

    def close(self):
        for eng in self._engines.values():
            eng.close()
        self._engines = {}
        self.__dict__.pop("_gather_cache", None)      # batch.likelihood_sharded's device / pinned buffers
