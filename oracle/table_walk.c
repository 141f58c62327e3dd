/*
 * TEST INFRASTRUCTURE ONLY - plain C walk through the packed tables: all four rsd models + likelihood.
 *
 * The same table-driven algebra the CUDA kernels run (reference lines below are
 * victor/ccf_model.py and victor/ccf_fit.py), written as straightforward scalar C with libm sqrt / exp /
 * divide and OpenMP over parameter rows.  It exists so that WHOLE bench batches (65 536 rows) can be
 * recomputed on the CPU in seconds and compared with the GPU row by row, and as an algorithm-for-algorithm
 * CPU timing next to the scipy oracle.  It reads the tables through the C ABI structs of
 * include/victor_b200.h; nothing in the product links or loads it (only tests/, smoke() and bench.py's
 * cpu_baseline leg do).  It is itself held to the scipy oracle and to the goldens of the unmodified
 * reference by tests/test_table_walk.py.
 *
 * Scope: everything the kernels cover -- rsd_model 'streaming', 'dispersion', 'kaiser' and 'euclid_special', up
 * to three real-space multipoles, sigma_v(r) and sigma_v(r, mu) templates, template or from-data coordinates,
 * every growth mode, the empirical velocity correction, and the chi-square / log-likelihood with all forms.
 *
 *   gcc -O2 -fopenmp -shared -fPIC -o oracle/_build/libtable_walk.so oracle/table_walk.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/victor_b200.h"

/* cell of u: bucket lookup plus bounded scan (tables.py: bucket_map); local coordinate clamped at 0 below
 * the first knot, where every spline is its boundary value (FITPACK ext=3) */
static int find_cell(const vb200_model_tables *m, double u, double *t) {
    double bf = floor(u * m->inv_h);
    int b = 0;
    if (bf >= 0.0) b = (bf < (double)(m->nbucket - 1)) ? (int)bf : m->nbucket - 1;
    int32_t e = m->bucket_base[b];
    int cell = (int)(e & 0x7fffffff);
    if (e < 0)
        for (int sc = 0; sc < m->maxscan; ++sc) cell += (u >= m->upper[cell]) ? 1 : 0;
    double tt = u - m->origin[cell];
    *t = (tt < 0.0) ? 0.0 : tt; /* NaN stays NaN */
    return cell;
}

static double cubic(const double *c, double t) { return ((c[3] * t + c[2]) * t + c[1]) * t + c[0]; }

static int beta_interval(const double *grid, int n, double b) {
    int k = 0;
    for (int i = 1; i < n - 1; ++i) k += (b >= grid[i]) ? 1 : 0;
    return k;
}

/* normalised dispersion template: the 1-D cubic, or the bicubic patch of a sigma_v(r, mu) template with mu
 * clamped to the template's range (RectBivariateSpline.ev -> bispeu clamps its arguments), :654-655 */
static double sv_at(const vb200_model_tables *m, int cell, double t, double mur) {
    if (m->sv_ny == 0) return cubic(m->sv + 4 * (size_t)cell, t);
    const double *yb = m->sv_ybreaks;
    double mc = mur;
    if (mc < yb[0]) mc = yb[0];
    if (mc > yb[m->sv_ny]) mc = yb[m->sv_ny];
    int yc = 0;
    for (int i = 1; i < m->sv_ny; ++i) yc += (mc >= yb[i]) ? 1 : 0;
    const double w = mc - yb[yc];
    const double *T = m->sv2d + ((size_t)cell * m->sv_ny + yc) * 16;
    double py[4];
    for (int q = 0; q < 4; ++q) py[q] = cubic(T + 4 * q, w);
    return cubic(py, t);
}

static double legendre_even(int ell, double x) {
    double x2 = x * x;
    if (ell == 0) return 1.0;
    if (ell == 2) return 1.5 * x2 - 0.5;
    return (4.375 * x2 - 3.75) * x2 + 0.375;
}

/* xi(s_j, mu_k) for one parameter row: xi[k * ns + j].  ccf_model.py:589-690 */
static void row_xi(const vb200_model_tables *m, const double *pr, const double *s, int ns, const double *mu, int nmu,
                   double *cells /* [3 + 2][ncell][4] scratch */, double *xi) {
    const int nc = m->ncell, per = nc * 4;
    const double fs8 = pr[0], sigv = pr[2], aperp = pr[3], apar = pr[4], astar = pr[5];
    const double beta = m->beta_dependent ? pr[1] : m->beta_fixed;
    const double eps = aperp / apar, iaHt = m->iaH * apar;
    double f;
    if (m->vel_indep_AP) {
        f = astar; /* :606-607 */
    } else {
        f = 0.0; /* :609-610 */
        for (int i = 0; i < m->nresc; ++i) {
            double mm = m->mu_resc[i];
            f += m->w_resc[i] * (apar * sqrt(1.0 + (1.0 - mm * mm) * (eps * eps - 1.0)));
        }
    }
    const double brow = pr[9];
    const int has_b = m->linear_bias && (brow == brow);
    const double bs = has_b ? m->bias / brow : 1.0;
    double Av;
    if (m->growth_mode == 2) {
        Av = fs8 / m->template_fsigma8 * m->growth_scale / apar; /* :439-443 */
    } else {
        double g = (m->growth_mode ? pr[1] * (has_b ? brow : m->bias) : fs8 / m->template_sigma8) * bs; /* :425-435 */
        Av = -g / (3.0 * iaHt);
    }
    const double B = Av / sigv, kappa = sigv * iaHt / f, Ae = pr[8] * bs;

    /* this row's cell cubics */
    int kb = 0;
    double tb = 0.0;
    if (m->beta_dependent) {
        kb = beta_interval(m->beta_grid, m->nbeta, beta);
        tb = beta - m->beta_grid[kb];
    }
    double *xi_c = cells, *v0_c = cells + 3 * (size_t)per, *d0_c = cells + 4 * (size_t)per;
    const double G = iaHt * Av / f, Mk = pr[6], Qk = pr[7];
    const size_t ell_stride = (size_t)(m->nbeta - 1) * 4 * per;
    for (int l = 0; l < m->n_ell; ++l) {
        const double *tab = m->xi_tab + l * ell_stride + (size_t)kb * 4 * per;
        for (int i = 0; i < per; ++i)
            xi_c[(size_t)l * per + i] = ((tab[3 * per + i] * tb + tab[2 * per + i]) * tb + tab[per + i]) * tb + tab[i];
    }
    for (int i = 0; i < per; ++i) {
        double v;
        if (m->vd_beta_dependent) {
            const double *tv = m->v0 + (size_t)kb * 4 * per;
            v = ((tv[3 * per + i] * tb + tv[2 * per + i]) * tb + tv[per + i]) * tb + tv[i];
        } else if (m->v0b) {
            v = m->v0[i] + Ae * m->v0b[i]; /* :451-455 */
        } else {
            v = m->v0[i];
        }
        v0_c[i] = v;
        double d;
        if (m->vd_beta_dependent) {
            const double *td = m->d0 + (size_t)kb * 4 * per;
            d = ((td[3 * per + i] * tb + td[2 * per + i]) * tb + td[per + i]) * tb + td[i];
        } else if (m->d0b) {
            d = m->d0[i] + Ae * m->d0b[i]; /* :456-459 */
        } else {
            d = m->d0[i];
        }
        d0_c[i] = d;
    }

    for (int k = 0; k < nmu; ++k) {
        const double sq = sqrt(1.0 - mu[k] * mu[k]);
        for (int j = 0; j < ns; ++j) {
            const double Sperp = s[j] * sq * (aperp / f), Spar = s[j] * mu[k] * (apar / f); /* :642-643 */
            const double Sp2 = Sperp * Sperp;
            double acc = 0.0, result;
            const double rt_data = s[j] * sq; /* s_perp / aperp in the fiducial cosmology, :676 */
#define XI_REAL(cell_, t_, mur_, rp_, out_)                                                                  \
    do {                                                                                                       \
        int c__ = (cell_);                                                                                     \
        double t__ = (t_), m__ = (mur_);                                                                       \
        if (m->realspace_from_data) { /* :675-679: back to fiducial coordinates */                             \
            const double rpd__ = (rp_) * f / apar;                                                             \
            const double rd__ = sqrt(rpd__ * rpd__ + rt_data * rt_data);                                       \
            c__ = find_cell(m, rd__, &t__);                                                                    \
            m__ = rpd__ / rd__;                                                                                \
        }                                                                                                      \
        (out_) = cubic(xi_c + 4 * (size_t)c__, t__); /* :683-687 */                                            \
        for (int l_ = 1; l_ < m->n_ell; ++l_)                                                                  \
            (out_) += cubic(xi_c + (size_t)l_ * per + 4 * (size_t)c__, t__) * legendre_even(m->ells[l_], m__); \
    } while (0)
            if (m->rsd_model == VB200_RSD_STREAMING) {
                for (int mi = 0; mi < m->nx; ++mi) {
                    const double xm = m->x[mi];
                    const double rp = Spar - xm * kappa;              /* :648 */
                    const double u = sqrt(Sp2 + rp * rp);              /* :651 */
                    const double mur = rp / u;                         /* :652 */
                    double t, xir;
                    const int cell = find_cell(m, u, &t);
                    const double sv = sv_at(m, cell, t, mur);             /* :654-655 */
                    const double z = (xm - B * cubic(v0_c + 4 * (size_t)cell, t) * mur) / sv; /* :656 */
                    XI_REAL(cell, t, mur, rp, xir);
                    acc += m->wx[mi] * (1.0 + xir) * exp(-0.5 * z * z) / sv; /* :690 */
                }
                result = acc - 1.0;
            } else if (m->rsd_model == VB200_RSD_DISPERSION) {       /* :659-671 */
                const double Strue = sqrt(Sp2 + Spar * Spar);
                double t0;
                const int c0 = find_cell(m, Strue, &t0);
                const double first = 1.0 + G * cubic(v0_c + 4 * (size_t)c0, t0) / Strue;
                for (int mi = 0; mi < m->nx; ++mi) {
                    const double xm = m->x[mi], num = Spar - xm * kappa;
                    double rp = num / first, u, t;
                    int cell;
                    for (int it = 0; it < m->niter; ++it) {
                        u = sqrt(Sp2 + rp * rp);
                        cell = find_cell(m, u, &t);
                        rp = num / (1.0 + G * cubic(v0_c + 4 * (size_t)cell, t) / u);
                    }
                    u = sqrt(Sp2 + rp * rp);
                    const double mur = rp / u;
                    cell = find_cell(m, u, &t);
                    const double sv = sv_at(m, cell, t, mur);
                    const double v0u = cubic(v0_c + 4 * (size_t)cell, t) / u;
                    const double jd = 1.0 + G * v0u + G * mur * mur * (cubic(d0_c + 4 * (size_t)cell, t) - v0u);
                    const double z = xm / sv;
                    double xir;
                    XI_REAL(cell, t, mur, rp, xir);
                    acc += m->wx[mi] * (1.0 + xir) * (1.0 / jd) * exp(-0.5 * z * z) / sv;
                }
                result = acc - 1.0;
            } else {                                                    /* kaiser / euclid_special :692-741 */
                const double MG = Mk * G;
                double rp = Spar, u, t;
                int cell;
                if (m->kaiser_coord_shift) {
                    const double Strue = sqrt(Sp2 + Spar * Spar);
                    cell = find_cell(m, Strue, &t);
                    rp = Spar / (1.0 + MG * cubic(v0_c + 4 * (size_t)cell, t) / Strue);
                    for (int it = 0; it < m->niter; ++it) {
                        u = sqrt(Sp2 + rp * rp);
                        cell = find_cell(m, u, &t);
                        rp = Spar / (1.0 + MG * cubic(v0_c + 4 * (size_t)cell, t) / u);
                    }
                }
                u = sqrt(Sp2 + rp * rp);
                const double mur = rp / u;
                cell = find_cell(m, u, &t);
                const double v0u = cubic(v0_c + 4 * (size_t)cell, t) / u;
                const int euclid = m->rsd_model == VB200_RSD_EUCLID;
                const double J = (euclid ? 3.0 : 1.0) * MG * v0u +
                                 (euclid ? 2.0 : 1.0) * MG * Qk * mur * mur * (cubic(d0_c + 4 * (size_t)cell, t) - v0u);
                double xir;
                XI_REAL(cell, t, mur, rp, xir);
                result = (euclid || m->kaiser_approximation) ? Mk * xir - J : (1.0 + Mk * xir) / (1.0 + J) - 1.0;
            }
#undef XI_REAL
            xi[(size_t)k * ns + j] = result;
        }
    }
}

static int supported(const vb200_model_tables *m) {
    return m->rsd_model >= VB200_RSD_STREAMING && m->rsd_model <= VB200_RSD_EUCLID && m->n_ell >= 1 &&
           m->n_ell <= VB200_MAX_POLES && m->sv_ny >= 0;
}

/* multipoles [n][L][ns] and / or xi [n][nmu][ns] on caller-supplied grids */
int tw_theory(const vb200_model_tables *m, const double *params, int64_t n, const double *s, int32_t ns,
              const double *mu, int32_t nmu, const double *wmu, int32_t L, double *xi_out, double *mult_out) {
    if (!supported(m)) return -4;
    int rc = 0;
#pragma omp parallel
    {
        double *cells = (double *)malloc(sizeof(double) * 5 * (size_t)m->ncell * 4);
        double *xi = (double *)malloc(sizeof(double) * (size_t)nmu * ns);
        if (!cells || !xi) rc = -3;
#pragma omp for schedule(dynamic, 4)
        for (int64_t r = 0; r < n; ++r) {
            if (rc) continue;
            row_xi(m, params + r * VB200_NPAR, s, ns, mu, nmu, cells, xi);
            if (xi_out) memcpy(xi_out + (size_t)r * nmu * ns, xi, sizeof(double) * (size_t)nmu * ns);
            if (mult_out)
                for (int l = 0; l < L; ++l)
                    for (int j = 0; j < ns; ++j) {
                        double acc = 0.0; /* :824-825, utils.py:45-56 */
                        for (int k = 0; k < nmu; ++k) acc += wmu[(size_t)l * nmu + k] * xi[(size_t)k * ns + j];
                        mult_out[((size_t)r * L + l) * ns + j] = acc;
                    }
        }
        free(cells);
        free(xi);
    }
    return rc;
}

/* chi2 and lnL of one row from its theory vector.  ccf_fit.py:166-260, 325-354, 441-483 */
static void row_like(const vb200_fit_tables *f, double beta, const double *th, double *res, double *chi2_out,
                     double *lnl_out) {
    const int p = f->ns * f->npoles;
    int kd = 0;
    double td = 0.0;
    if (f->data_beta_dependent) {
        kd = beta_interval(f->beta_ccf, f->nbeta_ccf, beta);
        td = beta - f->beta_ccf[kd];
    }
    const double *dt = f->data_tab + (size_t)kd * 4 * p;
    for (int j = 0; j < p; ++j) res[j] = th[j] - (((dt[3 * p + j] * td + dt[2 * p + j]) * td + dt[p + j]) * td + dt[j]);
    int lo = 0, hi = 0;
    double w = 0.0;
    if (!f->cov_fixed) {
        const int nb = f->nbeta_cov;
        const double *g = f->beta_cov;
        if (beta < g[0]) {
            lo = hi = 0;
        } else if (beta > g[nb - 1]) {
            lo = hi = nb - 1;
        } else {
            int below = 0, exact = -1;
            for (int i = 0; i < nb; ++i) {
                below += (g[i] < beta) ? 1 : 0;
                if (g[i] == beta) exact = i;
            }
            if (exact >= 0) {
                lo = hi = exact;
            } else if (below == 0) {
                w = beta; /* NaN */
            } else {
                lo = below - 1;
                hi = nb - 1; /* sic: last index with grid >= beta, ccf_fit.py:226, 258 */
                w = (beta - g[lo]) / (g[hi] - g[lo]);
            }
        }
    }
    double q[2] = {0.0, 0.0};
    for (int which = 0; which < ((hi != lo) ? 2 : 1); ++which) {
        const double *M = f->icov + (size_t)(which ? hi : lo) * p * p;
        double acc = 0.0;
        for (int i = 0; i < p; ++i) {
            double row = 0.0;
            for (int j = 0; j < p; ++j) row += M[(size_t)i * p + j] * res[j];
            acc += res[i] * row;
        }
        q[which] = acc;
    }
    double chi2 = (hi != lo) ? (1.0 - w) * q[0] + w * q[1] : ((w != w) ? w : q[0]);
    double norm = 0.0;
    if (f->use_logdet) {
        double ld = 0.0;
        if (hi != lo)
            for (int j = 0; j < p; ++j) ld += log1p(w * (f->lam[(size_t)lo * p + j] - 1.0));
        norm = -0.5 * (f->logdet[lo] + ld);
    }
    double lnl = (f->like_kind == VB200_LIKE_LOG) ? -f->like_a * log(1.0 + chi2 / f->like_nm1) / 2.0 + norm
                                                   : -0.5 * chi2 * f->like_a + norm;
    if (lnl != lnl) {
        lnl = -INFINITY;
        chi2 = INFINITY;
    }
    *chi2_out = chi2;
    *lnl_out = lnl;
}

/* theory [n][p] (or NULL), chi2 [n], lnl [n] on the fit's own grids */
int tw_likelihood(const vb200_model_tables *m, const vb200_fit_tables *f, const double *params, int64_t n,
                  double *theory, double *chi2, double *lnl) {
    if (!supported(m)) return -4;
    const int ns = f->ns, nmu = f->nmu, L = f->npoles, p = ns * L;
    int rc = 0;
#pragma omp parallel
    {
        double *cells = (double *)malloc(sizeof(double) * 5 * (size_t)m->ncell * 4);
        double *xi = (double *)malloc(sizeof(double) * (size_t)nmu * ns);
        double *th = (double *)malloc(sizeof(double) * 2 * (size_t)p);
        if (!cells || !xi || !th) rc = -3;
#pragma omp for schedule(dynamic, 4)
        for (int64_t r = 0; r < n; ++r) {
            if (rc) continue;
            const double *pr = params + r * VB200_NPAR;
            row_xi(m, pr, f->s, ns, f->mu, nmu, cells, xi);
            for (int l = 0; l < L; ++l)
                for (int j = 0; j < ns; ++j) {
                    double acc = 0.0;
                    for (int k = 0; k < nmu; ++k) acc += f->wmu[(size_t)l * nmu + k] * xi[(size_t)k * ns + j];
                    th[l * ns + j] = acc;
                }
            if (theory) memcpy(theory + (size_t)r * p, th, sizeof(double) * p);
            row_like(f, pr[1], th, th + p, chi2 + r, lnl + r);
        }
        free(cells);
        free(xi);
        free(th);
    }
    return rc;
}

int tw_abi_check(int64_t sizeof_model_tables, int64_t sizeof_fit_tables) {
    return (sizeof_model_tables == (int64_t)sizeof(vb200_model_tables) &&
            sizeof_fit_tables == (int64_t)sizeof(vb200_fit_tables))
               ? 0
               : -1;
}
