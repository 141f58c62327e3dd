/*
 * victor_b200.h - C ABI of the B200-native likelihood hot path of seshnadathur/victor.
 *
 * The reference is pure Python; it has no FFI of its own.  The entry points below are what a
 * binding for this path replaces, one call per public reference method:
 *
 *   vb200_create      <- CCFModel.__init__ / CCFFit.__init__ state hand-over
 *                        (victor/ccf_model.py:33-97, victor/ccf_fit.py:15-42): the host-built
 *                        spline / template / covariance tables are deep-copied to one GPU.
 *   vb200_theory      <- CCFModel.theory_xi            (victor/ccf_model.py:538-789)
 *                        CCFModel.theory_multipoles     (victor/ccf_model.py:791-827,
 *                                                        victor/utils.py:9-58)
 *                        CCFModel.theory_multipole_vector (victor/ccf_model.py:829-860)
 *                        on caller-supplied s / mu grids, for n parameter rows at once.
 *   vb200_theory_pairs <- CCFModel.theory_xi_2D        (victor/ccf_model.py:862-894): xi at
 *                        scattered (s, mu) points instead of an outer-product grid.
 *   vb200_likelihood  <- CCFFit.chi_squared / CCFFit.log_likelihood
 *                        (victor/ccf_fit.py:325-354, 356-483) and, through them,
 *                        CCFLikelihood.calculate (victor/likelihoods/CCFLikelihood.py:32-42),
 *                        on the data's own s grid, for n parameter rows at once.
 *
 * Conventions
 *   - All floating point is IEEE float64.  Arrays are C-contiguous, row-major.
 *   - Every function returns 0 on success or a negative VB200_E* code; the message is
 *     available from vb200_last_error() (thread-local).  Nothing throws across the ABI.
 *   - `params`, and every output pointer, may be a HOST pointer or a DEVICE pointer (memory of
 *     the context's GPU); the library detects which.  Host buffers are staged through device
 *     scratch with the copies inside the call, and the call returns after the results are in
 *     the host buffers.  With device pointers only, the call is asynchronous on `stream`
 *     (a cudaStream_t, or NULL for the legacy default stream) and the caller synchronises.
 *   - A context belongs to one GPU and is not re-entrant; distinct contexts are independent.
 *     Its scratch buffers (staging, theory vectors, the one-launch kernel's tickets) are shared by
 *     all of its calls: asynchronous calls on ONE context must go to one stream at a time (or be
 *     ordered by events); use one context per stream / host thread for concurrent work.
 *   - Per-row numerical failure (NaN anywhere in the row) gives lnlike = -inf, chi2 = +inf,
 *     as the reference does (victor/ccf_fit.py:477-481).
 */
#ifndef VICTOR_B200_H
#define VICTOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VB200_NPAR 10 /* columns of a parameter row, in this order: */
enum { VB200_P_FSIGMA8 = 0, VB200_P_BETA, VB200_P_SIGMA_V, VB200_P_APERP, VB200_P_APAR,
       VB200_P_ASTAR, VB200_P_M, VB200_P_Q,
       VB200_P_AV,   /* params.get('Av', 0): empirical correction of the mean velocity (ccf_model.py:453) */
       VB200_P_BIAS  /* params.get('bias', model['bias']) (ccf_model.py:359, 430); NaN = use the model's */ };

enum { VB200_OK = 0, VB200_EINVAL = -1, VB200_ECUDA = -2, VB200_ENOMEM = -3, VB200_EUNSUPPORTED = -4 };

enum { VB200_RSD_STREAMING = 0, VB200_RSD_DISPERSION = 1, VB200_RSD_KAISER = 2, VB200_RSD_EUCLID = 3 };
enum { VB200_LIKE_LINEAR = 0, VB200_LIKE_LOG = 1 };

#define VB200_MAX_POLES 3

/* Model half (CCFModel).  Radial tables live on one common cell set in the template
 * coordinate u = r / f: ncell cells, cell c covering [upper[c-1], upper[c]), cell 0 open to
 * the left, the last cell open to the right; `origin[c]` is the local-coordinate origin of
 * cell c; every cubic is c0 + c1 t + c2 t^2 + c3 t^3 with t = u - origin[cell].  A uniform
 * bucket grid (spacing 1 / inv_h, first bucket at u = 0) gives the first candidate cell;
 * at most `maxscan` upward steps follow. */
typedef struct vb200_model_tables {
    double iaH;               /* (1 + z_eff) / (100 E(z_eff)), ccf_model.py:45 */
    double template_sigma8;   /* ccf_model.py:187 */
    double beta_fixed;        /* beta used when the real-space input has no beta dependence (:585) */
    double inv_h;
    int32_t vel_indep_AP;     /* model['velocity_independent_of_AP'] (:606) */
    int32_t rsd_model;        /* VB200_RSD_* */
    int32_t n_ell;            /* real-space multipoles in xi(r, mu_r): 1 = assume_isotropic */
    int32_t ells[VB200_MAX_POLES];
    int32_t beta_dependent;   /* real-space input depends on reconstruction beta */
    int32_t ncell, nbucket, maxscan;
    int32_t nbeta;            /* length of beta_grid (>= 2) */
    int32_t nx;               /* velocity nodes */
    int32_t nresc;            /* nodes of the AP rescaling trapezoid (:609-610) */
    int32_t realspace_from_data;  /* model['realspace_ccf']['from_data'] (:675-679) */
    int32_t kaiser_approximation; /* (:737-741) */
    int32_t kaiser_coord_shift;   /* (:696-703) */
    int32_t niter;            /* fixed-point iterations of the dispersion / kaiser coordinate map (5) */
    int32_t sv_ny;            /* mu intervals of a sigma_v(r, mu) template; 0 = isotropic template */
    int32_t vd_beta_dependent; /* v0 / d0 are power tables in beta like xi_tab (matter model linear_bias) */
    int32_t growth_mode;      /* 0: growth = fsigma8 / template_sigma8; 1: beta * bias (:425-435);
                                 2: velocity template, v_r = v0(r) fsigma8 / template_fsigma8 * growth_scale / apar (:439-443, 484) */
    int32_t linear_bias;      /* matter model linear_bias: v0 / d0 carry 1 / bias, a per-row bias rescales them (:359-367) */
    double bias;              /* model['bias'] */
    double template_fsigma8;  /* velocity_pdf.mean.template_fsigma8 (:229), growth_mode 2 only */
    double growth_scale;      /* template_hubble_ratio (1 + z_sim) / (1 + z_eff) (:441-443), growth_mode 2 only */
    const double *origin;     /* [ncell] */
    const double *upper;      /* [ncell], last = +inf */
    const int32_t *bucket_base; /* [nbucket] */
    const double *beta_grid;  /* [nbeta] */
    const double *xi_tab;     /* [n_ell][nbeta-1][4 powers of (beta-beta_k)][ncell][4] */
    const double *v0;         /* [ncell][4]  spline of r * Delta(r)               (:449, 635) */
    const double *d0;         /* [ncell][4]  spline of 3 (delta - 2 Delta / 3)    (:450, 636) */
                              /* both [nbeta-1][4][ncell][4] when vd_beta_dependent;
                                 growth_mode 2: splines of the velocity template and of its finite-difference slope (:484-488) */
    const double *v0b;        /* [ncell][4] or NULL: empirical correction, v0 + Av v0b = spline of r Delta (1 + Av delta) (:454) */
    const double *d0b;        /* [ncell][4] or NULL: same for the slope term, d0 + Av d0b (:456-459) */
    const double *sv;         /* [ncell][4]  normalised sigma_v(r) template       (:654) */
    const double *sv2d;       /* [ncell][sv_ny][4][4] bicubic sigma_v(r, mu) patches, t^q w^p (NULL if sv_ny = 0) */
    const double *sv_ybreaks; /* [sv_ny + 1] mu breakpoints of sv2d */
    const double *x;          /* [nx] linspace(-6, 6) (:570) */
    const double *wx;         /* [nx] Simpson weights / sqrt(2 pi) (:690, :656) */
    const double *mu_resc;    /* [nresc] */
    const double *w_resc;     /* [nresc] trapezoid weights */
} vb200_model_tables;

/* Likelihood half (CCFFit); pass NULL to vb200_create for a model-only context. */
typedef struct vb200_fit_tables {
    int32_t ns;               /* length of the data s grid */
    int32_t npoles;           /* multipoles in the data vector; p = npoles * ns */
    int32_t nmu;              /* mu nodes of the projection (100) */
    int32_t data_beta_dependent;
    int32_t nbeta_ccf;        /* >= 2 */
    int32_t cov_fixed;
    int32_t nbeta_cov;        /* >= 1 */
    int32_t like_kind;        /* VB200_LIKE_* */
    int32_t use_logdet;
    double like_a;            /* LINEAR: lnL = -a chi2 / 2 + norm;  LOG: lnL = -a log(1 + chi2 / nm1) / 2 + norm */
    double like_nm1;
    const double *s;          /* [ns] */
    const double *mu;         /* [nmu] */
    const double *wmu;        /* [npoles][nmu] projection weights */
    const double *beta_ccf;   /* [nbeta_ccf] */
    const double *data_tab;   /* [nbeta_ccf-1][4][p] PCHIP power table of the data vector */
    const double *beta_cov;   /* [nbeta_cov] */
    const double *icov;       /* [nbeta_cov][p][p] */
    const double *logdet;     /* [nbeta_cov] log det of each covariance */
    const double *lam;        /* [nbeta_cov][p] generalised eigenvalues of (cov[last], cov[i]) */
} vb200_fit_tables;

typedef struct vb200_ctx vb200_ctx;

const char *vb200_version(void);
/* 0 when the caller's view of the two table structs has the library's sizes (binding self-check). */
int vb200_abi_check(int64_t sizeof_model_tables, int64_t sizeof_fit_tables);
const char *vb200_last_error(void);
int vb200_device_count(void);

int vb200_create(const vb200_model_tables *model, const vb200_fit_tables *fit, int device, vb200_ctx **out);
void vb200_destroy(vb200_ctx *ctx);

/* Kernel variant switches (integers; 0 selects the default where a default exists):
 *   "fast_math"  1 = hand-rolled rsqrt / rcp / exp (default), 0 = CUDA libm (the parity-test variant)
 *   "tuned"      1 = tuned kernels where they apply (default), 0 = always the general kernel
 *   "newton"     refinement of the MUFU seeds in the tuned streaming kernel: 2 = one Newton step (default), 3 = cubic step
 *   "exp_degree" 5 = degree-5 remainder polynomial on a 32-entry table (default), 3 = degree 3 on a 1024-entry table
 *   "ilp"        velocity nodes per loop trip: 0 = the kernel's default (10 streaming, 8 dispersion), 4 or 1 = the older
 *                variants (measurement)
 *   "threads"    block size of the batch kernels: 0 = automatic (128 tuned kernels, 256 general kernel) or 32..256;
 *                "nsplit" blocks per parameter row (0 = automatic)
 *   "fuse"       chi2 / lnL in the epilogue of the theory kernel when one block owns a row: 0 never, 1 where measured
 *                faster (default), 2 always
 *   "bucket"     chi2 kernel of batches of 4096 rows and more: rows grouped by covariance bracket first, both precision
 *                matrices of a group served from shared memory (1 = default; 0 = every row streams its own matrix)
 *   "tiny"       calls of one or two rows through the one-launch kernel k_small (1 = default)
 *   "mapped"     k_small writes its results into mapped page-locked host memory and the host polls a flag (1 = default)
 *   "graph"      replay calls of up to 256 host rows as one CUDA graph (1 = default)
 *   "chunks"     row chunks of host-bound bulk outputs (0 = automatic) */
int vb200_set_option(vb200_ctx *ctx, const char *key, int64_t value);

/* xi(s, mu) and / or its projections for n parameter rows.
 *   params [n][VB200_NPAR]; s [ns], mu [nmu], wmu [L][nmu] are HOST arrays (copied per call);
 *   xi_out [n][nmu][ns] or NULL; mult_out [n][L][ns] or NULL (requires wmu). */
int vb200_theory(vb200_ctx *ctx, const double *params, int64_t n,
                 const double *s, int32_t ns, const double *mu, int32_t nmu,
                 const double *wmu, int32_t L,
                 double *xi_out, double *mult_out, void *stream);

/* xi at npairs separate points (s[j], mu[j]) for n parameter rows -- what CCFModel.theory_xi_2D
 * (victor/ccf_model.py:862-894) gets from its 2500 scalar theory_xi calls, in one launch.
 *   s [npairs], mu [npairs] HOST arrays; xi_out [n][npairs]. */
int vb200_theory_pairs(vb200_ctx *ctx, const double *params, int64_t n,
                       const double *s, const double *mu, int32_t npairs,
                       double *xi_out, void *stream);

/* Theory vector, chi-square and log-likelihood for n parameter rows on the fit's own grids.
 *   theory [n][p] or NULL; chi2 [n] or NULL; lnlike [n] or NULL. */
int vb200_likelihood(vb200_ctx *ctx, const double *params, int64_t n,
                     double *theory, double *chi2, double *lnlike, void *stream);

int vb200_synchronize(vb200_ctx *ctx);

/* Kernel launches issued by this context so far (for bench.py's gpu_launches). */
int64_t vb200_launch_count(const vb200_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* VICTOR_B200_H */
