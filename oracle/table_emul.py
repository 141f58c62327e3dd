"""TEST INFRASTRUCTURE ONLY - numpy walk-through of the table-driven algorithm the kernels use.

Consumes the packed tables the product builds on the host (``victor_b200.tables``) and
evaluates them with plain numpy, batched over parameter rows.  It is the bridge between the
per-point scipy oracle (``oracle/ccf_oracle.py``, 0.13-0.25 s per point) and the CUDA kernels:
it checks the host tables and the restructured algebra on the CPU at sizes the scipy oracle
cannot reach, and it is the CPU reference for configurations the reference has no knob for
(dense mu / velocity grids).  Never imported by the product.

Algebra (reference lines in victor/ccf_model.py):
    u-units: everything radial is divided by the template rescaling factor f (:606-613)
    S_perp = s sqrt(1-mu^2) aperp / f,  S_par = s mu apar / f                    (:642-643)
    R_par = S_par - x_m sigma_v iaH apar / f ;  u = sqrt(S_perp^2 + R_par^2)      (:648-651)
    mu_r = R_par / u                                                              (:652)
    z = (x_m - (A_v / sigma_v) V0(u) mu_r) / SV(u),  A_v = -(fs8/s8_t) / (3 iaH apar)
    xi(s, mu) = sum_m wx_m (1 + xi_r(u)) exp(-z^2/2) / SV(u) - 1                  (:654-656, 690)
    xi_l(s) = sum_k W_l[k] xi(s, mu_k)                                            (:824-825)
"""
import numpy as np

from victor_b200 import tables as T


def _cells(mt, u):
    """Cell index and local coordinate for every u (bucket lookup + bounded scan)."""
    b = np.floor(u * mt.inv_h)
    b = np.clip(np.nan_to_num(b, nan=0.0), 0, len(mt.bucket_base) - 1).astype(np.int64)
    entry = mt.bucket_base[b].astype(np.int64)
    flagged = entry < 0
    cell = entry & 0x7FFFFFFF
    for _ in range(mt.maxscan):
        cell = cell + (flagged & (u >= mt.upper[cell]))
    return cell, np.where(u - mt.origin[cell] < 0, 0.0, u - mt.origin[cell])


def _horner(c, cell, t):
    k = c[cell]
    return ((k[..., 3] * t + k[..., 2]) * t + k[..., 1]) * t + k[..., 0]


def _beta_interval(grid, beta):
    k = np.searchsorted(grid, beta, side="right") - 1
    k = np.clip(k, 0, len(grid) - 2)
    return k, beta - grid[k]


def point_scalars(mt, rows):
    fs8, beta, sig, aperp, apar, astar = (rows[:, i] for i in range(6))
    eps = aperp / apar
    iaHt = mt.iaH * apar
    if mt.vel_indep_AP:
        f = astar.copy()
    else:
        y = apar[:, None] * np.sqrt(1 + (1 - mt.mu_resc[None, :] ** 2) * (eps[:, None] ** 2 - 1))
        f = (y * mt.w_resc[None, :]).sum(axis=1)
    # a bias given with the row rescales the linear_bias profiles (tabulated with the model's bias)
    brow = rows[:, 9] if rows.shape[1] > 9 else np.full(len(rows), np.nan)
    has_b = bool(mt.linear_bias) & ~np.isnan(brow)
    bs = np.where(has_b, mt.bias / np.where(has_b, brow, 1.0), 1.0)
    if mt.growth_mode == T.GROWTH_VELOCITY_TEMPLATE:
        Av = fs8 / mt.template_fsigma8 * mt.growth_scale / apar
    else:
        growth = beta * np.where(has_b, brow, mt.bias) if mt.growth_mode else fs8 / mt.template_sigma8
        Av = -(growth * bs) / (3 * iaHt)
    emp = (rows[:, 8] if rows.shape[1] > 8 else np.zeros(len(rows))) * bs
    return dict(eps=eps, iaHt=iaHt, f=f, Av=Av, emp=emp)


def xi_cells(mt, beta):
    """Per-row xi_l cell coefficients [n][n_ell][ncell][4] from the beta power table."""
    n = len(beta)
    if mt.beta_dependent:
        k, t = _beta_interval(mt.beta_grid, beta)
    else:
        k, t = np.zeros(n, dtype=np.int64), np.zeros(n)
    tab = mt.xi_tab[:, k]                       # [n_ell][n][4][ncell][4]
    tt = t[None, :, None, None]
    c = ((tab[:, :, 3] * tt + tab[:, :, 2]) * tt + tab[:, :, 1]) * tt + tab[:, :, 0]
    return np.moveaxis(c, 0, 1)                 # [n][n_ell][ncell][4]


def _sv(mt, cell, t, mur):
    """Normalised dispersion template: 1-D cubic, or the bicubic patch with mu clamped."""
    if mt.sv2d is None:
        return _horner(mt.sv, cell, t)
    yb = mt.sv_ybreaks
    mc = np.clip(mur, yb[0], yb[-1])
    yc = np.clip(np.searchsorted(yb, np.nan_to_num(mc, nan=yb[0]), side="right") - 1, 0, len(yb) - 2)
    w = mc - yb[yc]
    K = mt.sv2d[cell, yc]                       # [..., q, p]
    py = ((K[..., 3] * w[..., None] + K[..., 2]) * w[..., None] + K[..., 1]) * w[..., None] + K[..., 0]
    return ((py[..., 3] * t + py[..., 2]) * t + py[..., 1]) * t + py[..., 0]


def _beta_cells(mt, tab, beta):
    """Per-row [n][ncell][4] cubics from a [nbint][4][ncell][4] beta power table."""
    k, t = _beta_interval(mt.beta_grid, beta)
    T4 = tab[k]
    tt = t[:, None, None]
    return ((T4[:, 3] * tt + T4[:, 2]) * tt + T4[:, 1]) * tt + T4[:, 0]


def _legendre_even(ell, x):
    if ell == 0:
        return np.ones_like(x)
    if ell == 2:
        return 1.5 * x * x - 0.5
    return (4.375 * x * x - 3.75) * x * x + 0.375


def _xi_real(mt, xc, idx, cell, t, mur):
    """sum_l xi_l(u) L_l(mu_r) from the per-row cell coefficients xc [c][n_ell][ncell][4]."""
    tot = 0.0
    for li in range(mt.n_ell):
        kl = xc[idx, li, cell]
        xil = ((kl[..., 3] * t + kl[..., 2]) * t + kl[..., 1]) * t + kl[..., 0]
        tot = tot + (xil if li == 0 else xil * _legendre_even(int(mt.ells[li]), mur))
    return tot


def theory_xi(mt, rows, s, mu, chunk=32):
    """xi(s, mu) for each row: [n][nmu][ns].  All rsd models of the kernels (streaming, dispersion,
    kaiser, euclid_special), isotropic or anisotropic real-space input, model or from-data
    coordinates; isotropic template dispersion.  Mirrors victor_b200/csrc/k1_general.cuh."""
    rows = np.asarray(rows, float)
    n = len(rows)
    out = np.empty((n, len(mu), len(s)))
    sq = np.sqrt(1 - mu ** 2)
    rsd = mt.rsd_model
    x = mt.x if rsd in (T.RSD_STREAMING, T.RSD_DISPERSION) else np.zeros(1)
    for a in range(0, n, chunk):
        R = rows[a:a + chunk]
        sc = point_scalars(mt, R)
        sig = R[:, 2]
        beta = R[:, 1] if mt.beta_dependent else np.full(len(R), mt.beta_fixed)
        xc = xi_cells(mt, beta)                                    # [c][n_ell][ncell][4]
        idx = np.arange(len(R))[:, None, None, None]
        if mt.vd_beta_dependent:
            v0c, d0c = _beta_cells(mt, mt.v0, beta), _beta_cells(mt, mt.d0, beta)
        elif mt.v0b is not None:      # empirical correction (1 + Av delta) of the mean velocity
            v0c = mt.v0[None] + sc["emp"][:, None, None] * mt.v0b[None]
            d0c = mt.d0[None] + sc["emp"][:, None, None] * mt.d0b[None]
        else:
            v0c = np.broadcast_to(mt.v0, (len(R),) + mt.v0.shape)
            d0c = np.broadcast_to(mt.d0, (len(R),) + mt.d0.shape)

        def V0(cell, t):
            k = v0c[idx, cell]
            return ((k[..., 3] * t + k[..., 2]) * t + k[..., 1]) * t + k[..., 0]

        def D0(cell, t):
            k = d0c[idx, cell]
            return ((k[..., 3] * t + k[..., 2]) * t + k[..., 1]) * t + k[..., 0]

        f = sc["f"][:, None, None, None]
        apar = R[:, 4][:, None, None, None]
        Sperp = (s[None, None, :] * sq[None, :, None] * (R[:, 3] / sc["f"])[:, None, None])[..., None]
        Spar = (s[None, None, :] * mu[None, :, None] * (R[:, 4] / sc["f"])[:, None, None])[..., None]
        rt_data = (s[None, None, :] * sq[None, :, None] * np.ones(len(R))[:, None, None])[..., None]
        kap = (sig * sc["iaHt"] / sc["f"])[:, None, None, None]
        B = (sc["Av"] / sig)[:, None, None, None]
        G = (sc["iaHt"] * sc["Av"] / sc["f"])[:, None, None, None]
        xm = x[None, None, None, :]
        Sperp2 = Sperp ** 2

        def xi_at(rp, cell, t, mur):
            if not mt.from_data:
                return _xi_real(mt, xc, idx, cell, t, mur)
            rpd = rp * f / apar
            rd = np.sqrt(rpd ** 2 + rt_data ** 2)
            cd, td = _cells(mt, rd + 0 * rp)
            return _xi_real(mt, xc, idx, cd, td, rpd / rd)

        with np.errstate(invalid="ignore", divide="ignore"):
            if rsd == T.RSD_STREAMING:
                rp = Spar - xm * kap
                u = np.sqrt(Sperp2 + rp ** 2)
                mur = rp / u
                cell, t = _cells(mt, u)
                svv = _sv(mt, cell, t, mur)
                z = (xm - B * V0(cell, t) * mur) / svv
                integrand = (1 + xi_at(rp, cell, t, mur)) * np.exp(-0.5 * z * z) / svv
                out[a:a + chunk] = (integrand * mt.wx[None, None, None, :]).sum(axis=-1) - 1
            elif rsd == T.RSD_DISPERSION:
                Strue = np.sqrt(Sperp2 + Spar ** 2)
                c0, t0 = _cells(mt, Strue)
                num = Spar - xm * kap
                rp = num / (1 + G * V0(c0, t0) / Strue)
                for _ in range(mt.niter):
                    u = np.sqrt(Sperp2 + rp ** 2)
                    cell, t = _cells(mt, u)
                    rp = num / (1 + G * V0(cell, t) / u)
                u = np.sqrt(Sperp2 + rp ** 2)
                mur = rp / u
                cell, t = _cells(mt, u)
                svv = _sv(mt, cell, t, mur)
                v0u = V0(cell, t) / u
                jac = 1 / (1 + G * v0u + G * mur ** 2 * (D0(cell, t) - v0u))
                z = xm / svv
                integrand = (1 + xi_at(rp, cell, t, mur)) * jac * np.exp(-0.5 * z * z) / svv
                out[a:a + chunk] = (integrand * mt.wx[None, None, None, :]).sum(axis=-1) - 1
            else:
                Mk = R[:, 6][:, None, None, None]
                Qk = R[:, 7][:, None, None, None]
                MG = Mk * G
                rp = Spar + 0 * MG
                if mt.kaiser_coord_shift:
                    Strue = np.sqrt(Sperp2 + Spar ** 2)
                    c0, t0 = _cells(mt, Strue)
                    rp = Spar / (1 + MG * V0(c0, t0) / Strue)
                    for _ in range(mt.niter):
                        u = np.sqrt(Sperp2 + rp ** 2)
                        cell, t = _cells(mt, u)
                        rp = Spar / (1 + MG * V0(cell, t) / u)
                u = np.sqrt(Sperp2 + rp ** 2)
                mur = rp / u
                cell, t = _cells(mt, u)
                v0u = V0(cell, t) / u
                ca, cb = (3.0, 2.0) if rsd == T.RSD_EUCLID else (1.0, 1.0)
                J = ca * MG * v0u + cb * MG * Qk * mur ** 2 * (D0(cell, t) - v0u)
                xi = xi_at(rp, cell, t, mur)
                if rsd == T.RSD_EUCLID or mt.kaiser_approximation:
                    res = Mk * xi - J
                else:
                    res = (1 + Mk * xi) / (1 + J) - 1
                out[a:a + chunk] = res[..., 0]
    return out


def theory_multipoles(mt, rows, s, mu, W, chunk=32):
    """[n][L][ns] multipoles and [n][nmu][ns] xi."""
    xi = theory_xi(mt, rows, s, mu, chunk)
    return np.einsum("lk,nkj->nlj", W, xi), xi


def chi2_lnl(ft, beta, theory):
    """chi2[n], lnL[n] from stacked theory vectors [n][p].  victor/ccf_fit.py:166-260, 325-483."""
    n = len(beta)
    if ft.data_beta_dependent:
        k, t = _beta_interval(ft.beta_ccf, beta)
    else:
        k, t = np.zeros(n, dtype=np.int64), np.zeros(n)
    tab = ft.data_tab[k]                                            # [n][4][p]
    tt = t[:, None]
    d = ((tab[:, 3] * tt + tab[:, 2]) * tt + tab[:, 1]) * tt + tab[:, 0]
    resid = theory - d
    chi2 = np.empty(n)
    lnl = np.empty(n)
    g = ft.beta_cov
    for i in range(n):
        if ft.cov_fixed:
            lo, hi, w = 0, 0, 0.0
        elif beta[i] < g.min():
            lo, hi, w = 0, 0, 0.0
        elif beta[i] > g.max():
            lo, hi, w = len(g) - 1, len(g) - 1, 0.0
        elif beta[i] in g:
            lo = hi = int(np.where(g == beta[i])[0][0])
            w = 0.0
        elif np.isnan(beta[i]):
            chi2[i], lnl[i] = np.inf, -np.inf
            continue
        else:
            lo = int(np.where(g < beta[i])[0][-1])
            hi = len(g) - 1                                         # the reference's bracket quirk
            w = (beta[i] - g[lo]) / (g[hi] - g[lo])
        qlo = resid[i] @ ft.icov[lo] @ resid[i]
        qhi = resid[i] @ ft.icov[hi] @ resid[i] if hi != lo else qlo
        chi2[i] = (1 - w) * qlo + w * qhi
        norm = 0.0
        if ft.use_logdet:
            ld = ft.logdet[lo] + (np.log1p(w * (ft.lam[lo] - 1)).sum() if hi != lo else 0.0)
            norm = -0.5 * ld
        if ft.like_kind == T.LIKE_LOG:
            lnl[i] = -ft.like_a * np.log(1 + chi2[i] / ft.like_nm1) / 2 + norm
        else:
            lnl[i] = -0.5 * chi2[i] * ft.like_a + norm
        if np.isnan(lnl[i]):
            lnl[i], chi2[i] = -np.inf, np.inf
    return chi2, lnl
