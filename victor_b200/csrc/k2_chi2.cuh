// K2: chi-square and log-likelihood, one warp per parameter row.  Replaces
//   CCFFit.multipole_datavector / get_interpolated_{covariance,precision}
//                                         victor/ccf_fit.py:166-260, 306-323
//   CCFFit.chi_squared                    victor/ccf_fit.py:349-354
//   CCFFit.log_likelihood                 victor/ccf_fit.py:441-483
#pragma once
#include "common.cuh"

namespace vb200 {

constexpr int kK2Warps = 8;
constexpr int kK2MaxChunks = 8;  // p <= 256

__device__ __forceinline__ double quad_form(const double *M, const double *res, int p, int lane) {
    // y_j = sum_i M[i][j] res_i with lanes over columns j (coalesced rows; M symmetric in exact
    // arithmetic, and res^T M res does not depend on which index is contracted first)
    double y[kK2MaxChunks];
#pragma unroll
    for (int c = 0; c < kK2MaxChunks; ++c) y[c] = 0.0;
    for (int i = 0; i < p; ++i) {
        const double ri = res[i];
        const double *rowp = M + (size_t)i * p;
#pragma unroll
        for (int c = 0; c < kK2MaxChunks; ++c) {
            const int j = lane + 32 * c;
            if (j < p) y[c] = fma(rowp[j], ri, y[c]);
        }
    }
    double q = 0.0;
#pragma unroll
    for (int c = 0; c < kK2MaxChunks; ++c) {
        const int j = lane + 32 * c;
        if (j < p) q = fma(y[c], res[j], q);
    }
    return warp_sum(q);
}

__global__ void __launch_bounds__(kK2Warps * 32) k_chi2(const __grid_constant__ K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FitDev &f = a.f;
    const int p = f.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *res = reinterpret_cast<double *>(smem_raw) + (size_t)warp * p;
    const long long row = (long long)blockIdx.x * kK2Warps + warp;
    if (row >= a.n) return;
    const double beta = a.params[row * kNPar + 1];
    const double *th = a.theory + (size_t)row * p;

    // residual against the PCHIP-in-beta data vector (ccf_fit.py:193, 322-323, 350)
    int kd = 0;
    double td = 0.0;
    if (f.data_beta_dependent) {
        kd = beta_interval(f.beta_ccf, f.nbeta_ccf, beta);
        td = beta - f.beta_ccf[kd];
    }
    const double *dt = f.data_tab + (size_t)kd * 4 * p;
    for (int j = lane; j < p; j += 32) {
        const double d = fma(fma(fma(dt[3 * p + j], td, dt[2 * p + j]), td, dt[p + j]), td, dt[j]);
        res[j] = th[j] - d;
    }
    __syncwarp();

    // matrix bracket with the reference's conventions (ccf_fit.py:218-227, 250-259)
    int lo = 0, hi = 0;
    double w = 0.0;
    if (!f.cov_fixed) {
        const int nb = f.nbeta_cov;
        const double *g = f.beta_cov;
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
            } else if (below == 0) {  // beta is NaN: every comparison false
                lo = hi = 0;
                w = beta;
            } else {
                lo = below - 1;
                hi = nb - 1;  // sic: last index with grid >= beta
                w = (beta - g[lo]) / (g[hi] - g[lo]);
            }
        }
    }
    const double qlo = quad_form(f.icov + (size_t)lo * p * p, res, p, lane);
    double chi2 = qlo;
    if (hi != lo) {
        const double qhi = quad_form(f.icov + (size_t)hi * p * p, res, p, lane);
        chi2 = (1.0 - w) * qlo + w * qhi;
    } else if (w != w) {
        chi2 = w;
    }

    double norm = 0.0;
    if (f.use_logdet) {
        double ld = 0.0;
        if (hi != lo) {
            const double *lam = f.lam + (size_t)lo * p;
            for (int j = lane; j < p; j += 32) ld += log1p(w * (lam[j] - 1.0));
            ld = warp_sum(ld);
        }
        norm = -0.5 * (f.logdet[lo] + ld);
    }
    if (lane == 0) {
        double lnl;
        if (f.like_kind == 1)
            lnl = -f.like_a * log(1.0 + chi2 / f.like_nm1) / 2.0 + norm;  // ccf_fit.py:457, 469
        else
            lnl = -0.5 * chi2 * f.like_a + norm;                          // ccf_fit.py:462, 471
        if (lnl != lnl) {  // ccf_fit.py:477-481
            lnl = -INFINITY;
            chi2 = INFINITY;
        }
        if (a.chi2) a.chi2[row] = chi2;
        if (a.lnl) a.lnl[row] = lnl;
    }
}
}  // namespace vb200
