"""Host-side table builder: everything the CUDA kernels read, built once with scipy.

The reference re-fits FITPACK / PCHIP splines inside every likelihood call
(victor/ccf_model.py:299-326, 615-636, 654; victor/ccf_fit.py:193).  All of those fits are
linear in their data and their abscissae only get rescaled by one per-point factor, so here
they are done ONCE on the host, with the same scipy routines, and flattened into
piecewise-polynomial coefficient tables the device evaluates directly:

* every radial spline (xi_l(u; beta), V0(u), D0(u), sigma_v template SV(u)) is re-expressed
  on ONE set of cells -- the union of all their knots in the template coordinate u = r/f --
  so the kernel does a single cell search and a single local coordinate per quadrature point;
  outside a spline's own knot range its cell polynomial is the constant boundary value, which
  is exactly FITPACK's ``ext=3`` / ``bispeu`` clamping;
* the beta dependence of xi^r and of the data vector (PCHIP over the reconstruction grid,
  extrapolating outside it) becomes, per beta interval, a cubic in (beta - beta_k) whose four
  coefficient sets are themselves spline fits (fit of the PCHIP coefficient vectors);
* the mu-projection (bicubic interp2d -> 200-point trapezoid with Legendre weights,
  ccf_model.py:824-825, utils.py:45-56) and the Simpson rule in velocity (ccf_model.py:690)
  become fixed weight vectors;
* the log-determinant of the (linearly blended) covariance uses the generalised eigenvalues
  of (cov[hi], cov[lo]):  logdet((1-t) C_lo + t C_hi) = logdet C_lo + sum log1p(t (lam - 1)).
"""
import functools
from dataclasses import dataclass, field

import numpy as np
from scipy.integrate import simpson
from scipy.interpolate import InterpolatedUnivariateSpline, PchipInterpolator, PPoly
from scipy.special import legendre

from .utils import InputError, trapezoid

RSD_STREAMING, RSD_DISPERSION, RSD_KAISER, RSD_EUCLID = 0, 1, 2, 3
LIKE_LINEAR, LIKE_LOG = 0, 1          # lnL = -a/2 chi2 + norm   |   lnL = -m/2 log(1 + chi2/(nm-1)) + norm
MAX_POLES = 3
NPAR = 10                              # fsigma8, beta, sigma_v, aperp, apar, astar, M, Q, Av, bias
PARAM_ORDER = ("fsigma8", "beta", "sigma_v", "aperp", "apar", "astar", "M", "Q", "Av", "bias")
GROWTH_FSIGMA8, GROWTH_BETA_BIAS, GROWTH_VELOCITY_TEMPLATE = 0, 1, 2


# --------------------------------------------------------------------------------------------
# quadrature weight vectors
# --------------------------------------------------------------------------------------------
def velocity_nodes(nx=50):
    """x nodes and weights of the velocity integral.

    ``x = linspace(-6, 6)`` (ccf_model.py:570) and ``simps(..., x=v_par)`` (:690) with
    v_par = x sigma_v.  Returns (x, w) with  integral f dv  =  sigma_v * sum_m w[m] f_m, the
    weights being whatever scipy's composite Simpson rule assigns on this grid (for an even
    number of points that includes its asymmetric end correction).
    """
    if not 3 <= nx <= 128:
        raise InputError("velocity_nodes must be between 3 and 128")
    x = np.linspace(-6, 6, nx)
    w = simpson(np.eye(nx), x=x, axis=1)
    return x, w


def mu_nodes(poles, nmu=100):
    """mu grid of theory_multipoles (ccf_model.py:816-822): [0,1] for even poles, else [-1,1]."""
    poles = np.atleast_1d(poles)
    even = not np.any(poles % 2)
    return (np.linspace(0, 1, nmu) if even else np.linspace(-1, 1, nmu)), even


def mu_projection_weights(poles, nmu=100, npts=200):
    """(mu, W) -- see _mu_projection_weights; cached per (poles, nmu, npts): building the nmu unit splines takes
    9 ms for 100 nodes, which would otherwise dominate every single-point theory_multipoles call."""
    mu, W = _mu_projection_weights(tuple(int(p) for p in np.atleast_1d(poles)), int(nmu), int(npts))
    return mu.copy(), W.copy()


@functools.lru_cache(maxsize=64)
def _mu_projection_weights(poles, nmu, npts):
    """Weights W[l, k] with  xi_l(s_j) = sum_k W[l, k] xi(s_j, mu_k).

    The reference builds ``interp2d(s, mu, xi, kind='cubic')`` and integrates it over 200 mu
    values at each s_j with the trapezoid rule against (2l+1) L_l (ccf_model.py:824-825,
    utils.py:45-56).  At a data abscissa s_j the tensor-product interpolant reduces to the
    1-D not-a-knot cubic spline in mu through the row xi(s_j, :), so the whole operation is a
    fixed linear functional of that row, independent of the s grid.
    """
    poles = np.atleast_1d(poles)
    mu, even = mu_nodes(poles, nmu)
    fine = np.linspace(0.0, 1.0, npts) if even else np.linspace(-1, 1, npts)
    basis = np.empty((nmu, npts))
    eye = np.eye(nmu)
    for k in range(nmu):
        basis[k] = InterpolatedUnivariateSpline(mu, eye[k], k=3)(fine)
    tw = np.empty(npts)
    d = np.diff(fine)
    tw[0], tw[-1] = d[0] / 2, d[-1] / 2
    tw[1:-1] = (d[:-1] + d[1:]) / 2
    W = np.empty((len(poles), nmu))
    for i, ell in enumerate(poles):
        pref = (2 * ell + 1) if even else (2 * ell + 1) / 2
        W[i] = pref * basis @ (tw * legendre(ell)(fine))
    return mu, W


# --------------------------------------------------------------------------------------------
# splines -> per-cell cubic coefficients on a common cell set
# --------------------------------------------------------------------------------------------
def _shift_cubic(c_desc, delta):
    """Re-expand p(d) = c3 d^3 + c2 d^2 + c1 d + c0 about d = delta; ascending output."""
    c3, c2, c1, c0 = (np.longdouble(v) for v in c_desc)
    dl = np.longdouble(delta)
    n0 = ((c3 * dl + c2) * dl + c1) * dl + c0
    n1 = (3 * c3 * dl + 2 * c2) * dl + c1
    n2 = 3 * c3 * dl + c2
    return np.array([n0, n1, n2, c3], dtype=np.float64)


def bspline_cells(tck, knots, lo, hi):
    """Cubic coefficients (ascending, local coordinate t = u - cell_origin) on every cell of
    ``knots`` of the cubic B-spline ``tck`` clamped to its boundary values outside [lo, hi]
    (FITPACK ``ext=3`` / the argument clamping of ``bispeu``).

    Cells: 0 = (-inf, knots[0]), i = [knots[i-1], knots[i]), last = [knots[-1], inf).
    Cell origins: knots[0] for cell 0, knots[i-1] otherwise.  The interior knots of the spline
    must be a subset of ``knots`` so that no cell straddles one of them.
    """
    from scipy.interpolate import BSpline
    t, c, k = tck
    n = len(t) - k - 1
    bs = BSpline(t, np.asarray(c, float)[:n], k, extrapolate=False)
    pp = PPoly.from_spline((t, np.asarray(c, float), k))
    brk, coef = pp.x, pp.c                      # coef[:, i] on [brk[i], brk[i+1]]
    ncell = len(knots) + 1
    out = np.zeros((ncell, 4))
    lo_val, hi_val = float(bs(lo)), float(bs(hi))
    for cidx in range(ncell):
        a = knots[0] if cidx == 0 else knots[cidx - 1]
        if cidx == 0 or a < lo:
            # whole cell below the spline's range?  (cell [a, b) with b <= lo)
            b = knots[0] if cidx == 0 else (knots[cidx] if cidx < len(knots) else np.inf)
            if b <= lo:
                out[cidx, 0] = lo_val
                continue
            raise InputError("spline_cells: cell straddles the first knot of a spline")
        if a >= hi:
            out[cidx, 0] = hi_val
            continue
        b = knots[cidx] if cidx < len(knots) else np.inf
        mid = 0.5 * (a + min(b, hi))
        i = int(np.searchsorted(brk, mid, side="right") - 1)
        i = min(max(i, 0), coef.shape[1] - 1)
        while brk[i + 1] <= brk[i]:             # skip zero-length end intervals
            i += 1
        out[cidx] = _shift_cubic(coef[:, i], a - brk[i])
    return out


def spline_cells(x, y, knots):
    """Per-cell cubics of ``InterpolatedUnivariateSpline(x, y, ext=3)``, see bspline_cells."""
    x = np.asarray(x, float)
    spl = InterpolatedUnivariateSpline(x, y, k=3)
    return bspline_cells(spl._eval_args, knots, x[0], x[-1])


def sv2d_cells(r_sv, mu_sv, sv_rmu, knots):
    """Bicubic patches of ``RectBivariateSpline(r_sv, mu_sv, sv_rmu.T)`` (ccf_model.py:654, 667)
    on (radial cell of ``knots``) x (mu interval): T[cell][ycell][q][p] multiplies t^q w^p with
    t = u - cell origin, w = mu - ybreaks[ycell].  ``.ev`` clamps both arguments to the knot
    range (``bispeu``), which is the boundary-constant rule in u and an explicit clamp in mu."""
    from scipy.interpolate import RectBivariateSpline
    r_sv, mu_sv = np.asarray(r_sv, float), np.asarray(mu_sv, float)
    spl = RectBivariateSpline(r_sv, mu_sv, np.asarray(sv_rmu, float).T)
    tx, ty, c = spl.tck
    nxb, nyb = len(tx) - 4, len(ty) - 4
    C = np.asarray(c, float).reshape(nxb, nyb)
    ncell = len(knots) + 1
    A = np.empty((ncell, 4, nyb))
    for j in range(nyb):
        A[:, :, j] = bspline_cells((tx, np.concatenate([C[:, j], np.zeros(4)]), 3), knots, r_sv[0], r_sv[-1])
    ybreaks = np.unique(ty)
    nyc = len(ybreaks) - 1
    T = np.zeros((ncell, nyc, 4, 4))
    for cell in range(ncell):
        for q in range(4):
            pp = PPoly.from_spline((ty, np.concatenate([A[cell, q], np.zeros(4)]), 3))
            keep = np.nonzero(np.diff(pp.x) > 0)[0]
            assert len(keep) == nyc
            T[cell, :, q, :] = pp.c[::-1, keep].T
    return np.ascontiguousarray(T), ybreaks


def union_knots(*grids):
    ks = np.unique(np.concatenate([np.asarray(g, float) for g in grids]))
    return ks


BUCKET_FLAG = np.int32(-2 ** 31)   # bit 31 of a bucket entry: a knot lies strictly inside the bucket


def _bucket_try(knots, inv_h):
    span = knots[-1]
    nb = int(np.floor(span * inv_h)) + 2
    h = 1.0 / inv_h
    starts = np.arange(nb) * h
    ends = starts + h
    ends[-1] = np.inf
    base = np.searchsorted(knots, starts, side="right")      # knots <= start: cell holding the start
    top = np.searchsorted(knots, ends, side="left")          # knots <  end : highest cell reachable
    # cell 0 (below the first knot) is never entered: the kernel clamps the local coordinate at
    # 0, and every spline below the first knot equals the first cell's cubic at t = 0 (ext=3)
    return np.maximum(base, 1), np.maximum(top, 1)


def bucket_map(knots, candidates=(512, 1024, 2048)):
    """Uniform bucket grid over [0, knots[-1]] -> first candidate cell, flag and scan length.

    Bucket b covers [b/inv_h, (b+1)/inv_h); the last one extends to +inf.  entry[b] = index of
    the cell containing the bucket start, with bit 31 set when further knots lie strictly inside
    the bucket, in which case the kernel advances while u >= upper(cell), at most ``maxscan``
    times.  The spacing is chosen among the knot lattice (1 / smallest gap, snapped to an exact
    reciprocal when it is one) and a few fine uniform grids, whichever flags the fewest buckets:
    knots on a lattice (BOSS: integers) give buckets that never straddle a knot.
    """
    knots = np.asarray(knots, float)
    span = knots[-1]
    lattice = 1.0 / np.diff(knots).min()
    if abs(lattice - round(lattice)) < 1e-9 * max(1.0, lattice):
        lattice = float(round(lattice))
    options = [lattice] + [(n - 2) / span for n in candidates]
    best = None
    for inv_h in options:
        if span * inv_h > 8190:
            continue
        base, top = _bucket_try(knots, inv_h)
        flagged = np.count_nonzero(top > base)
        score = (flagged / len(base), len(base))
        if best is None or score < best[0]:
            best = (score, inv_h, base, top)
    _, inv_h, base, top = best
    entry = base.astype(np.int32)
    entry[top > base] |= BUCKET_FLAG
    maxscan = int(np.max(top - base))
    return float(inv_h), entry, maxscan


# --------------------------------------------------------------------------------------------
# the packed tables
# --------------------------------------------------------------------------------------------
@dataclass
class ModelTables:
    """Parameter-independent tables of the model half (CCFModel)."""
    iaH: float
    template_sigma8: float
    vel_indep_AP: bool
    rsd_model: int
    n_ell: int                      # real-space multipoles entering xi(r, mu_r): 1 if isotropic
    ells: np.ndarray                # their orders
    beta_dependent: bool
    beta_fixed: float               # beta used when the input has no beta dependence
    knots: np.ndarray               # union knots [nk]; ncell = nk + 1
    origin: np.ndarray              # cell origins [ncell]
    upper: np.ndarray               # upper knot of each cell [ncell] (+inf for the last)
    inv_h: float
    bucket_base: np.ndarray         # int32 [nb]
    maxscan: int
    beta_grid: np.ndarray           # [nbeta] (>= 2) real-space reconstruction grid
    xi_tab: np.ndarray              # [n_ell][nbint][4 (power of t, ascending)][ncell][4]
    v0: np.ndarray                  # [ncell][4]
    d0: np.ndarray                  # [ncell][4]
    sv: np.ndarray                  # [ncell][4]   (isotropic sigma_v template)
    x: np.ndarray                   # [nx]
    wx: np.ndarray                  # [nx]  Simpson weights / sqrt(2 pi)
    mu_resc: np.ndarray             # [50] nodes of the AP rescaling trapezoid
    w_resc: np.ndarray              # [50] its weights
    vd_beta_dependent: bool = False # v0 / d0 are [nbint][4][ncell][4] power tables in beta (linear_bias)
    growth_mode: int = 0            # 0: fsigma8 / template_sigma8;  1: beta * bias (:429-430);  2: velocity template
    bias: float = 1.9
    linear_bias: bool = False       # v0 / d0 carry 1 / bias: a bias given with the parameter row rescales them
    template_fsigma8: float = 0.0   # growth_mode 2: v_r = V0(r) fsigma8 / template_fsigma8 growth_scale / apar (:439-443)
    growth_scale: float = 1.0       # template_hubble_ratio (1 + z_sim) / (1 + z_eff)
    v0b: np.ndarray = None          # [ncell][4] empirical correction: V0 = v0 + Av v0b, D0 = d0 + Av d0b (:451-459)
    d0b: np.ndarray = None
    sv2d: np.ndarray = None         # [ncell][nyc][4][4] bicubic sigma_v(u, mu) patches, or None (isotropic)
    sv_ybreaks: np.ndarray = None   # [nyc + 1] mu breakpoints of sv2d
    from_data: bool = False         # real-space ccf measured from data: xi at (r_par/apar, s_perp/aperp)
    kaiser_approximation: bool = False
    kaiser_coord_shift: bool = True
    niter: int = 5                  # fixed-point iterations of the dispersion / kaiser coordinate map
    extras: dict = field(default_factory=dict)

    @property
    def ncell(self):
        return len(self.knots) + 1


@dataclass
class FitTables:
    """Tables of the likelihood half (CCFFit)."""
    p: int
    data_beta_dependent: bool
    beta_ccf: np.ndarray            # [nbd] (>= 2)
    data_tab: np.ndarray            # [nbint_d][4][p]
    cov_fixed: bool
    beta_cov: np.ndarray            # [nbc]
    icov: np.ndarray                # [nbc][p][p]
    logdet: np.ndarray              # [nbc]
    lam: np.ndarray                 # [nbc][p]  generalised eigenvalues of (cov[last], cov[i])
    like_kind: int
    like_a: float                   # LINEAR: a;  LOG: m
    like_nm1: float                 # LOG: nmocks - 1
    use_logdet: bool


def linear_bias_maps(r, r_eval):
    """Matrices taking the real-space monopole values at ``r`` to  xi_0(r_eval)  and to
    3 / r_eval^3 * trapz_{100}(xi_0 r'^2)  (ccf_model.py:362-369; divide by the bias for delta, Delta)."""
    r = np.asarray(r, float)
    Ld = np.empty((len(r_eval), len(r)))
    LD = np.empty_like(Ld)
    eye = np.eye(len(r))
    for j in range(len(r)):
        xir = InterpolatedUnivariateSpline(r, eye[j], ext=3)
        Ld[:, j] = xir(r_eval)
        for i, ri in enumerate(r_eval):
            rr = np.linspace(0, ri, 100)
            LD[i, j] = 3 * trapezoid(xir(rr) * rr ** 2, rr) / ri ** 3
    return Ld, LD


def pchip_power_table(grid, values):
    """PCHIP over ``grid`` of ``values`` [ngrid][...] -> (coef[nint][4 ascending][...]).

    scipy's ``PchipInterpolator.c`` has shape (4 descending, nint, ...); with
    ``extrapolate=True`` (default; ccf_model.py:326, ccf_fit.py:193) the end polynomials are
    used outside the grid, so the interval index is simply clamped.
    """
    pc = PchipInterpolator(grid, values, axis=0)
    c = np.asarray(pc.c)                      # (4, nint, ...)
    return np.ascontiguousarray(np.moveaxis(c[::-1], 0, 1))  # (nint, 4 ascending, ...)


def build_model_tables(state, options, nx=50):
    """``state``: a loaded victor_b200.model.CCFModel; ``options``: its merged model dict."""
    if options["matter_model"] not in ("template", "linear_bias"):
        raise NotImplementedError(
            f"matter_model '{options['matter_model']}' has no B200 path (only 'template' and 'linear_bias')")
    if options["mean_model"] not in ("linear", "template"):
        raise NotImplementedError(f"mean-velocity model '{options['mean_model']}' has no B200 path "
                                  "(only 'linear' and 'template')")
    rsd = {"streaming": RSD_STREAMING, "dispersion": RSD_DISPERSION, "kaiser": RSD_KAISER,
           "euclid_special": RSD_EUCLID}.get(options["rsd_model"])
    if rsd is None:
        raise InputError(f"theory_xi: Unrecognised choice of model {options['rsd_model']}")

    r = np.asarray(state.r, float)
    r31 = np.append([0.01], r)
    knots = union_knots(r31, state.r_for_sv)
    ncell = len(knots) + 1
    origin = np.concatenate([[knots[0]], knots])
    upper = np.concatenate([knots, [np.inf]])
    inv_h, base, maxscan = bucket_map(knots)

    beta_dep = not state.fixed_real_input
    vd_beta_dep, growth_mode, bias = False, GROWTH_FSIGMA8, float(options.get("bias", 1.9))
    linear_bias = options["matter_model"] == "linear_bias"
    empirical = bool(options["empirical_corr"]) and options["mean_model"] == "linear"
    template_fsigma8, growth_scale = 0.0, 1.0
    v0b = d0b = None
    r_fine = np.linspace(0.1, r.max(), 100)      # "finer grid to better estimate derivative numerically" (:456, :486)

    def slope_at_r31(values_on_fine_grid):
        # dvr_interp = _spline(rgrid, np.gradient(vr_grid, rgrid), ext=3); dvr = dvr_interp(r)   (:457-459, 487-488)
        return InterpolatedUnivariateSpline(r_fine, np.gradient(values_on_fine_grid, r_fine), ext=3)(r31)

    def velocity_cells(d31, D31):
        """(v0, d0, v0b, d0b) cell cubics from delta(r31), Delta(r31): v_r = A_v (v0 + Av v0b)(r),
        dv_r/dr = A_v (d0 + Av d0b)(r) with A_v = -growth / (3 iaH)   (ccf_model.py:446-459)."""
        if not empirical:
            # velocity_terms() re-splines the profiles at their own abscissae before use (:422-423);
            # an interpolating spline evaluated at its knots returns the data, so V0/D0 data are:
            return (spline_cells(r31, r31 * D31, knots),
                    spline_cells(r31, 3.0 * (d31 - 2.0 * D31 / 3.0), knots), None, None)
        dS = InterpolatedUnivariateSpline(r31, d31, ext=3)          # :422-423
        DS = InterpolatedUnivariateSpline(r31, D31, ext=3)
        base, corr = r_fine * DS(r_fine), r_fine * DS(r_fine) * dS(r_fine)
        return (spline_cells(r31, r31 * D31, knots), spline_cells(r31, slope_at_r31(base), knots),
                spline_cells(r31, r31 * D31 * d31, knots), spline_cells(r31, slope_at_r31(corr), knots))

    if options["mean_model"] == "template":
        # velocity template (a testing option of the reference, ccf_model.py:227-246, 439-443, 483-488):
        # v_r = radial_velocity(r) * growth', growth' = fsigma8 / template_fsigma8 * hubble ratio * (1+z_sim)/(1+z_eff) / apar
        if not getattr(state, "has_velocity_template", False):
            raise InputError("velocity_terms: Cannot use template option as no template has been supplied.")
        growth_mode = GROWTH_VELOCITY_TEMPLATE
        template_fsigma8 = float(state.template_fsigma8)
        growth_scale = float(state.template_hubble_ratio) * (1 + state.z_sim) / (1 + state.z_eff)
        v0 = spline_cells(r31, state.radial_velocity(r31), knots)
        d0 = spline_cells(r31, slope_at_r31(state.radial_velocity(r_fine)), knots)
    elif options["matter_model"] == "template":
        # velocity templates (ccf_model.py:421-423, 449-450, 635-636): data at r31, knots r31
        v0, d0, v0b, d0b = velocity_cells(state.delta(r31), state.integrated_delta(r31))
    else:
        # linear_bias (ccf_model.py:358-370): delta = xi_0 / b, Delta(r) = 3 / (b r^3) times a 100-point
        # trapezoid of xi_0 r'^2 -- both linear in the monopole values, which are PCHIP cubics in beta:
        # the V0 / D0 cell cubics are therefore, per beta interval, cubics in (beta - beta_k) too
        Ld, LD = linear_bias_maps(r, r31)
        if options["realspace_ccf_from_data"]:
            growth_mode = GROWTH_BETA_BIAS        # growth term beta * bias (:429-430)
        mono = np.asarray(state.real_multipoles["0"], float)
        if beta_dep:
            if empirical:
                raise NotImplementedError("empirical_corr with a beta-dependent linear_bias matter model has no "
                                          "B200 path (delta * Delta is not a cubic in beta)")
            vd_beta_dep = True
            pw = pchip_power_table(np.asarray(state.beta, float), mono)      # (nbint, 4, nr)
            v0 = np.zeros((pw.shape[0], 4, ncell, 4))
            d0 = np.zeros_like(v0)
            for k in range(pw.shape[0]):
                for q in range(4):
                    D31, d31 = LD @ pw[k, q] / bias, Ld @ pw[k, q] / bias
                    v0[k, q] = spline_cells(r31, r31 * D31, knots)
                    d0[k, q] = spline_cells(r31, 3.0 * (d31 - 2.0 * D31 / 3.0), knots)
        else:
            v0, d0, v0b, d0b = velocity_cells(Ld @ mono / bias, LD @ mono / bias)

    # sigma_v template: all mu rows identical -> 1-D cubic spline in u (FITPACK tensor spline of a
    # function constant in mu is that 1-D spline)
    sv = spline_cells(state.r_for_sv, state.sv_rmu[0], knots)
    sv2d = sv_ybreaks = None
    if not state.sv_isotropic:
        # sigma_v(r, mu) template: true bicubic evaluation with FITPACK's argument clamping
        sv2d, sv_ybreaks = sv2d_cells(state.r_for_sv, state.mu_for_sv, state.sv_rmu, knots)

    # real-space multipoles
    iso = bool(options["assume_isotropic"])
    ells = np.array([0]) if iso else np.asarray(state.poles_r)
    n_ell = len(ells)
    if beta_dep:
        beta_grid = np.asarray(state.beta, float)
        nbint = len(beta_grid) - 1
        xi_tab = np.zeros((n_ell, nbint, 4, ncell, 4))
        for i, ell in enumerate(ells):
            pw = pchip_power_table(beta_grid, state.real_multipoles[f"{ell}"])  # (nbint, 4, nr)
            for k in range(nbint):
                for q in range(4):
                    xi_tab[i, k, q] = spline_cells(r, pw[k, q], knots)
    else:
        beta_grid = np.array([0.0, 1.0])
        xi_tab = np.zeros((n_ell, 1, 4, ncell, 4))
        for i, ell in enumerate(ells):
            xi_tab[i, 0, 0] = spline_cells(r, state.real_multipoles[f"{ell}"], knots)

    x, w = velocity_nodes(nx)
    mu_resc = np.linspace(1e-10, 1)
    d = np.diff(mu_resc)
    w_resc = np.zeros_like(mu_resc)
    w_resc[:-1] += d / 2
    w_resc[1:] += d / 2

    return ModelTables(
        iaH=float(state.iaH), template_sigma8=float(state.template_sigma8 or 1.0),
        vel_indep_AP=bool(options["velocity_independent_of_AP"]), rsd_model=rsd,
        n_ell=n_ell, ells=ells.astype(np.int32), beta_dependent=beta_dep, beta_fixed=0.40,
        knots=knots, origin=origin, upper=upper, inv_h=inv_h,
        bucket_base=base, maxscan=maxscan, beta_grid=beta_grid,
        xi_tab=np.ascontiguousarray(xi_tab), v0=v0, d0=d0, sv=sv,
        x=x, wx=w / np.sqrt(2 * np.pi), mu_resc=mu_resc, w_resc=w_resc,
        vd_beta_dependent=vd_beta_dep, growth_mode=growth_mode, bias=bias, linear_bias=linear_bias,
        template_fsigma8=template_fsigma8, growth_scale=growth_scale, v0b=v0b, d0b=d0b,
        sv2d=sv2d, sv_ybreaks=sv_ybreaks, from_data=bool(options["realspace_ccf_from_data"]),
        kaiser_approximation=bool(options.get("kaiser_approximation", False)),
        kaiser_coord_shift=bool(options.get("kaiser_coord_shift", True)),
        niter=int(options.get("niter", 5)))    # model.get('niter', 5), ccf_model.py:661, 701, 752


def likelihood_constants(like, p):
    """(kind, a, nm1) of the likelihood form.  ccf_fit.py:455-473."""
    form = str(like["form"]).lower()
    nm = like.get("nmocks", 1)
    if form == "sellentin":
        return LIKE_LOG, float(nm), float(nm - 1)
    if form == "hartlap":
        return LIKE_LINEAR, float((nm - p - 2) / (nm - 1)), 0.0
    if form == "percival":
        npar = like["nparams"]  # KeyError if absent, like the reference
        B = (nm - p - 2) / ((nm - p - 1) * (nm - p - 4))
        m = npar + 2 + (nm - 1 + B * (p - npar)) / (1 + B * (p - npar))
        return LIKE_LOG, float(m), float(nm - 1)
    if form == "gaussian":
        return LIKE_LINEAR, 1.0, 0.0
    raise InputError("Unrecognised likelihood form")


def build_fit_tables(fit, like):
    """``fit``: a loaded victor_b200.fit.CCFFit; ``like``: the likelihood option dict."""
    from scipy.linalg import eigh
    p = len(fit.s) * len(fit.poles_s)
    stack = np.array([fit.redshift_multipoles[f"{ell}"] for ell in fit.poles_s])
    if fit.fixed_data:
        beta_ccf = np.array([0.0, 1.0])
        data_tab = np.zeros((1, 4, p))
        data_tab[0, 0] = stack.reshape(p)
    else:
        beta_ccf = np.asarray(fit.beta_ccf, float)
        vals = np.moveaxis(stack, 1, 0).reshape(len(beta_ccf), p)   # [nbeta][l*ns + j]
        data_tab = pchip_power_table(beta_ccf, vals)
    if fit.fixed_covmat:
        cov = np.asarray(fit.covmat, float)[None]
        icov = np.asarray(fit.icov, float)[None]
        beta_cov = np.array([0.0])
    else:
        cov = np.asarray(fit.covmat, float)
        icov = np.asarray(fit.icov, float)
        beta_cov = np.asarray(fit.beta_covmat, float)
    nbc = len(cov)
    logdet = np.zeros(nbc)
    lam = np.ones((nbc, p))
    use_logdet = not fit.fixed_covmat
    if use_logdet:
        # The reference takes slogdet of the blended covariance per evaluation and returns (-inf, inf) for a point
        # whose matrix is not positive definite (ccf_fit.py:445-450).  A grid matrix that fails the check gets NaN
        # here: every row whose bracket uses it then ends in the NaN guard, (-inf, +inf); the other rows are served.
        good = np.zeros(nbc, dtype=bool)
        for i in range(nbc):
            sign, ld = np.linalg.slogdet(cov[i])
            try:
                np.linalg.cholesky(cov[i])
            except np.linalg.LinAlgError:
                sign = 0
            good[i] = sign == 1
            logdet[i] = ld if good[i] else np.nan
        for i in range(nbc - 1):
            if good[i] and good[-1]:
                lam[i] = eigh(cov[-1], cov[i], eigvals_only=True)
            else:
                lam[i] = np.nan
        if not good.all():
            from .utils import log
            log.warning("covariance matrices %s of the beta grid are not positive definite: parameter points that "
                        "use them evaluate to (-inf, inf)", np.flatnonzero(~good).tolist())
    kind, a, nm1 = likelihood_constants(like, p)
    return FitTables(p=p, data_beta_dependent=not fit.fixed_data, beta_ccf=beta_ccf,
                     data_tab=np.ascontiguousarray(data_tab), cov_fixed=bool(fit.fixed_covmat),
                     beta_cov=beta_cov, icov=np.ascontiguousarray(icov), logdet=logdet, lam=lam,
                     like_kind=kind, like_a=a, like_nm1=nm1, use_logdet=use_logdet)
