// K2: chi-square and log-likelihood.  Replaces
//   CCFFit.multipole_datavector / get_interpolated_{covariance,precision}
//                                         victor/ccf_fit.py:166-260, 306-323
//   CCFFit.chi_squared                    victor/ccf_fit.py:349-354
//   CCFFit.log_likelihood                 victor/ccf_fit.py:441-483
// Two forms of the same arithmetic:
//   k_chi2      one warp per parameter row, theory vectors read from global memory (after a K1 launch
//               that split rows over several blocks, or when the caller asked for the theory vectors);
//   block_chi2  epilogue of the K1 kernels when one block owns a whole row (batch mode): the theory
//               vector is still in shared memory, the block's warps split the rows of the precision
//               matrices, and neither a second launch nor the theory round trip through HBM is needed.
#pragma once
#include "common.cuh"

namespace vb200 {

constexpr int kK2Warps = 8;
constexpr int kK2MaxChunks = 8;  // p <= 256

// res^T M res as kK2Warps partial sums, partial `part` taking the matrix rows i = part, part + kK2Warps, ...:
//   sum_i sum_j M[i][j] res_i res_j  with lanes over columns j (coalesced rows; M symmetric in exact
// arithmetic, and res^T M res does not depend on which index is contracted first).  The partition is
// the same whether one warp walks through all the partials (k_chi2) or the warps of a block take one
// each (block_chi2), so both give bit-identical chi-squares.
// kC = number of 32-column chunks actually present (p <= 32 kC): the loops below run over kC chunks only -- with the
// data vector's p = 60 that is 2 of the kK2MaxChunks = 8 the general form walks through with predicates
template <int kC>
__device__ __forceinline__ double quad_partial_t(const double *M, const double *res, int p, int lane, int part) {
    double y[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) y[c] = 0.0;
    for (int i = part; i < p; i += kK2Warps) {
        const double ri = res[i];
        const double *rowp = M + (size_t)i * p;
#pragma unroll
        for (int c = 0; c < kC; ++c) {
            const int j = lane + 32 * c;
            if (j < p) y[c] = fma(rowp[j], ri, y[c]);
        }
    }
    double q = 0.0;
#pragma unroll
    for (int c = 0; c < kC; ++c) {
        const int j = lane + 32 * c;
        if (j < p) q = fma(y[c], res[j], q);
    }
    return warp_sum(q);
}

__device__ __forceinline__ double quad_partial(const double *M, const double *res, int p, int lane, int part) {
    switch ((p + 31) >> 5) {   // same sums in the same order whichever instantiation runs
        case 1: return quad_partial_t<1>(M, res, p, lane, part);
        case 2: return quad_partial_t<2>(M, res, p, lane, part);
        case 3: return quad_partial_t<3>(M, res, p, lane, part);
        case 4: return quad_partial_t<4>(M, res, p, lane, part);
        default: return quad_partial_t<kK2MaxChunks>(M, res, p, lane, part);
    }
}

// The same partial sums for FOUR residual vectors against one matrix (k_chi2_bucketed): every matrix element is fetched
// once and used four times.  `res4` holds the four vectors interleaved ([i][4]: two 16-byte broadcast loads per matrix
// row), `own[r][c]` is res_r[lane + 32 c] in registers.  Each row's operations and their order are those of
// quad_partial_t / quad_form: bit-identical results.
// shared-memory loads of data written earlier in the same kernel (ordered after the barrier by the memory clobber;
// the plain lds_f64 of common.cuh is for tables that never change once the kernel's main loop runs)
__device__ __forceinline__ double2 lds_f64x2_ordered(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double lds_f64_ordered(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}

template <int kC>
__device__ __forceinline__ void quad_form4_t(unsigned m_s, unsigned res_s, const double (&own)[4][kC], int p, int lane,
                                             double (&q)[4]) {
    // m_s / res_s: shared-window addresses of the matrix and of the interleaved residuals: the loop below advances two
    // 32-bit addresses and nothing else
    bool has[kC];
#pragma unroll
    for (int c = 0; c < kC; ++c) has[c] = lane + 32 * c < p;
#pragma unroll
    for (int r = 0; r < 4; ++r) q[r] = 0.0;
    const unsigned rowstep = 8u * (unsigned)p * kK2Warps;
    for (int part = 0; part < kK2Warps; ++part) {
        double y[4][kC];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < kC; ++c) y[r][c] = 0.0;
        unsigned ma = m_s + 8u * ((unsigned)part * p + lane), ra = res_s + 32u * part;
#pragma unroll 4
        for (int i = part; i < p; i += kK2Warps, ma += rowstep, ra += 32u * kK2Warps) {
            const double2 r01 = lds_f64x2_ordered(ra), r23 = lds_f64x2_ordered(ra + 16);
#pragma unroll
            for (int c = 0; c < kC; ++c) {
                if (has[c]) {
                    const double m = lds_f64_ordered(ma + 256u * c);
                    y[0][c] = fma(m, r01.x, y[0][c]);
                    y[1][c] = fma(m, r01.y, y[1][c]);
                    y[2][c] = fma(m, r23.x, y[2][c]);
                    y[3][c] = fma(m, r23.y, y[3][c]);
                }
            }
        }
        double t[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            t[r] = 0.0;
#pragma unroll
            for (int c = 0; c < kC; ++c)
                if (has[c]) t[r] = fma(y[r][c], own[r][c], t[r]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {   // four butterflies side by side (each one is warp_sum's)
#pragma unroll
            for (int r = 0; r < 4; ++r) t[r] += __shfl_xor_sync(0xffffffffu, t[r], o);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) q[r] += t[r];
    }
}

__device__ __forceinline__ double quad_form(const double *M, const double *res, int p, int lane) {
    double q = 0.0;
    for (int part = 0; part < kK2Warps; ++part) q += quad_partial(M, res, p, lane, part);
    return q;
}

// interval k with grid[k] <= b < grid[k+1], clamped (beta_interval of common.cuh) with the grid values spread
// over the lanes of a warp: one round of loads instead of a serial scan (n <= 33)
__device__ __forceinline__ int beta_interval_warp(const double *grid, int n, double b, int lane) {
    const bool ge = (lane >= 1 && lane < n - 1) ? (b >= grid[lane]) : false;
    return __popc(__ballot_sync(0xffffffffu, ge));
}

// data vector at beta: PCHIP power table (ccf_fit.py:193, 322-323, 350)
struct DataAt {
    const double *dt;
    double td;
    int p;
    __device__ __forceinline__ DataAt(const FitDev &f, double beta) : p(f.p) {
        int kd = 0;
        td = 0.0;
        if (f.data_beta_dependent) {
            kd = beta_interval(f.beta_ccf, f.nbeta_ccf, beta);
            td = beta - f.beta_ccf[kd];
        }
        dt = f.data_tab + (size_t)kd * 4 * p;
    }
    // the same with the interval found by the whole warp (beta_interval_warp: same interval, one round of loads)
    __device__ __forceinline__ DataAt(const FitDev &f, double beta, int lane) : p(f.p) {
        int kd = 0;
        td = 0.0;
        if (f.data_beta_dependent) {
            kd = f.nbeta_ccf <= 33 ? beta_interval_warp(f.beta_ccf, f.nbeta_ccf, beta, lane)
                                   : beta_interval(f.beta_ccf, f.nbeta_ccf, beta);
            td = beta - f.beta_ccf[kd];
        }
        dt = f.data_tab + (size_t)kd * 4 * p;
    }
    __device__ __forceinline__ double operator()(int j) const {
        return fma(fma(fma(dt[3 * p + j], td, dt[2 * p + j]), td, dt[p + j]), td, dt[j]);
    }
};

// matrix bracket with the reference's conventions (ccf_fit.py:218-227, 250-259): lower neighbour and
// the LAST grid index; end matrices outside the grid; the grid matrix itself on a grid value.
// A NaN beta comes back as w = NaN with lo == hi.
__device__ __forceinline__ void cov_bracket(const FitDev &f, double beta, int &lo, int &hi, double &w) {
    lo = hi = 0;
    w = 0.0;
    if (f.cov_fixed) return;
    const int nb = f.nbeta_cov;
    const double *g = f.beta_cov;
    if (beta < g[0]) return;
    if (beta > g[nb - 1]) {
        lo = hi = nb - 1;
        return;
    }
    int below = 0, exact = -1;
    for (int i = 0; i < nb; ++i) {
        below += (g[i] < beta) ? 1 : 0;
        if (g[i] == beta) exact = i;
    }
    if (exact >= 0) {
        lo = hi = exact;
    } else if (below == 0) {  // beta is NaN: every comparison false
        w = beta;
    } else {
        lo = below - 1;
        hi = nb - 1;  // sic: last index with grid >= beta
        w = (beta - g[lo]) / (g[hi] - g[lo]);
    }
}

// cov_bracket for a row whose lower bracket `lo` is already known (the bucketed kernel: lo is the tile's bracket):
// the remaining case analysis of cov_bracket on three grid values -- same hi, same w
__device__ __forceinline__ void bracket_from_lo(const FitDev &f, double beta, int lo, double g0, double glo, double glast,
                                                int &hi, double &w) {
    hi = lo;
    w = 0.0;
    if (f.cov_fixed) return;
    if (beta != beta) {   // NaN: cov_bracket leaves lo = hi = 0 and hands the NaN on
        w = beta;
        return;
    }
    if (beta < g0 || beta > glast || glo == beta) return;   // outside the grid: the end matrix; on a grid value: that matrix
    hi = f.nbeta_cov - 1;
    w = (beta - glo) / (glast - glo);
}

__device__ __forceinline__ double blend_chi2(double qlo, double qhi, int lo, int hi, double w) {
    if (hi != lo) return fma(1.0 - w, qlo, w * qhi);   // written out: every kernel form must round this the same way
    return (w != w) ? w : qlo;
}

// log det of the blended covariance from the generalised eigenvalues (tables.py: lam), one warp
__device__ __forceinline__ double norm_term(const FitDev &f, int lo, int hi, double w, int lane) {
    if (!f.use_logdet) return 0.0;
    double ld = 0.0;
    if (hi != lo) {
        const double *lam = f.lam + (size_t)lo * f.p;
        for (int j = lane; j < f.p; j += 32) ld += log1p(w * (lam[j] - 1.0));
        ld = warp_sum(ld);
    }
    return -0.5 * (f.logdet[lo] + ld);
}

__device__ __forceinline__ void store_likelihood(const FitDev &f, double chi2, double norm, long long row,
                                                 double *chi2_out, double *lnl_out) {
    double lnl;
    if (f.like_kind == 1)
        lnl = -f.like_a * log(1.0 + chi2 / f.like_nm1) / 2.0 + norm;  // ccf_fit.py:457, 469
    else
        lnl = -0.5 * chi2 * f.like_a + norm;                          // ccf_fit.py:462, 471
    if (lnl != lnl) {  // ccf_fit.py:477-481
        lnl = -INFINITY;
        chi2 = INFINITY;
    }
    if (chi2_out) chi2_out[row] = chi2;
    if (lnl_out) lnl_out[row] = lnl;
}

// Rows per warp of k_chi2.  The upper bracket of (almost) every row is the LAST precision matrix (the reference's
// bracket quirk, ccf_fit.py:225-227), so a block keeps that matrix in shared memory and serves all its
// kK2Warps * kK2RowsPerWarp rows from it: L2 traffic per row drops from two matrices to a little over one.
constexpr int kK2RowsPerWarp = 2;
constexpr int kK2StageMaxP = 64;   // p x p doubles must fit beside the residuals in the default 48 KB

__host__ __device__ inline bool k2_stages(int p) { return p <= kK2StageMaxP; }
__host__ __device__ inline size_t k2_smem_bytes(int p) {
    return ((size_t)kK2Warps * p + (k2_stages(p) ? (size_t)p * p : 0)) * sizeof(double);
}

// Bucketed form: rows sorted by lower bracket (k_bracket_count / k_bracket_scatter), then a block takes kK2TileRows
// rows of ONE bracket and keeps both of their precision matrices (and the bracket's eigenvalue row) in shared memory.
constexpr int kK2TileRows = 32;    // rows per block: 4 per warp, taken together
__host__ __device__ inline size_t k2_bucket_smem_bytes(int p) {
    const size_t pe = (size_t)((p + 1) & ~1), pp = ((size_t)p * p + 1) & ~(size_t)1;
    return (2 * pp + pe + (size_t)kK2Warps * 4 * pe) * sizeof(double);
}

// doubles of shared memory the fused epilogue needs: the theory / residual vector and the per-warp
// partial quadratic forms
__host__ __device__ inline int fused_fit_doubles(int p) { return ((p + 1) & ~1) + 2 * kK2Warps; }

// Fused epilogue: `th` [p] holds this row's theory vector in shared memory (written by
// write_outputs, visible after a __syncthreads); `red` has room for 2 * kK2Warps doubles.
__device__ __forceinline__ void block_chi2(const FitDev &f, double beta, double *th, double *red, long long row,
                                           double *chi2_out, double *lnl_out, int tid, int nthr) {
    const int p = f.p;
    const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
    const DataAt data(f, beta);
    for (int j = tid; j < p; j += nthr) th[j] -= data(j);
    __syncthreads();
    int lo, hi;
    double w;
    cov_bracket(f, beta, lo, hi, w);
    for (int part = warp; part < kK2Warps; part += nwarp) {   // (blocks of fewer than 8 warps take several)
        const double qlo = quad_partial(f.icov + (size_t)lo * p * p, th, p, lane, part);
        const double qhi = (hi != lo) ? quad_partial(f.icov + (size_t)hi * p * p, th, p, lane, part) : 0.0;
        if (lane == 0) {
            red[part] = qlo;
            red[kK2Warps + part] = qhi;
        }
    }
    __syncthreads();
    if (warp == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < kK2Warps; ++i) {   // same order as quad_form
            a += red[i];
            b += red[kK2Warps + i];
        }
        const double chi2 = blend_chi2(a, b, lo, hi, w);
        const double norm = norm_term(f, lo, hi, w, lane);
        if (lane == 0) store_likelihood(f, chi2, norm, row, chi2_out, lnl_out);
    }
}

}  // namespace vb200
