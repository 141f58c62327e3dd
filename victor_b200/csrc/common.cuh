// Shared definitions of the victor_b200 CUDA kernels (sm_100a): device-side table views,
// kernel argument structs and the hand-rolled FP64 math.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace vb200 {

constexpr int kMaxPoles = 3;
constexpr int kExpTab = 32;
constexpr int kExpTabBig = 1024;  // 2^(j/1024): exp variants 3 / 30 (degree-3 remainder polynomial)
constexpr int kMaxNx = 128;       // velocity nodes that fit in the kernel-parameter table
constexpr int kBucketFlag = (int)0x80000000;
constexpr int kNPar = 10;         // doubles per parameter row (VB200_NPAR)
constexpr int kNScal = 10;        // per-row scalars in shared memory (row_scalars_to_shared)

enum { kRsdStreaming = 0, kRsdDispersion = 1, kRsdKaiser = 2, kRsdEuclid = 3 };

struct ModelDev {
    double iaH, s8t, beta_fixed, inv_h;
    int vel_indep_AP, rsd_model, n_ell, beta_dependent;
    int ncell, nbucket, maxscan, nbeta, nx, nresc;
    int from_data, kaiser_approx, kaiser_shift, niter;
    int ells[kMaxPoles];
    const double *origin, *upper;
    const int *bucket_base;
    const double *beta_grid, *xi_tab, *v0, *d0, *sv, *x, *wx, *mu_resc, *w_resc;
    const double *exp_tab;  // [kExpTab] 2^(j/32)
    const double *exp_tab_big;  // [kExpTabBig] 2^(j/1024)
    const double *sv2d;     // [ncell][sv_ny][4][4] bicubic sigma_v(u, mu) patches (sv_ny > 0), else null
    const double *sv_yb;    // [sv_ny + 1] mu breakpoints
    int sv_ny;
    int vd_beta_dep, growth_mode;   // matter model linear_bias: v0 / d0 beta power tables; growth = beta * bias
    double bias;
    int lin_bias;           // v0 / d0 carry 1 / bias (matter model linear_bias): a per-row bias rescales them
    double fs8t, growth_scale;      // growth_mode 2 (velocity template): A_v = fsigma8 / fs8t * growth_scale / apar
    const double *v0b, *d0b;        // [ncell][4] empirical-correction parts: V0 = v0 + Av v0b (null if unused)
};

struct FitDev {
    int p, data_beta_dependent, nbeta_ccf, cov_fixed, nbeta_cov, like_kind, use_logdet;
    double like_a, like_nm1;
    const double *beta_ccf, *data_tab, *beta_cov, *icov, *logdet, *lam;
};

struct K1Args {
    ModelDev m;
    const double *params;
    long long n;
    const double *s, *mu, *sqmu, *wmu;  // [ns], [nmu], [nmu] sqrt(1-mu^2), [L][nmu]
    int ns, nmu, L;
    int jper, nsplit;
    int has_flags;     // some bucket entry carries the interior-knot flag (general kernel: comparison scan needed)
    int pairwise;      // 1: nmu == 1 and mu / sqmu have ns entries -- xi at the pairs (s_j, mu_j) (theory_xi_2D)
    double *xi_out;    // [n][nmu][ns] or null
    double *mult_out;  // [n][L][ns]  or null
    double xw[2 * kMaxNx];  // x_m then Simpson weight / sqrt(2 pi): read through the constant bank
    // fused likelihood epilogue (nsplit == 1 only): the block that produced a row's theory vector also
    // contracts it with the precision matrices (k2_chi2.cuh: block_chi2), so K2 is not launched and the
    // theory vector need not leave the SM
    int fuse;
    FitDev f;
    double *chi2, *lnl;   // [n] each, either may be null
};

constexpr int kK2MaxBins = 64;   // brackets the bucketed K2 can tell apart (nbeta_cov <= 64)

struct K2Args {
    FitDev f;
    const double *params;
    const double *theory;  // [n][p]
    long long n;
    double *chi2, *lnl;    // either may be null
    // bucketed form (k2_chi2.cuh: k_chi2_bucketed): rows grouped by their lower covariance bracket
    int *order;            // [n] row indices, bucket after bucket; null = not bucketed
    unsigned char *lo8;    // [n] lower bracket of every row
    unsigned *bins;        // [2 * kK2MaxBins]: rows per bracket | scatter cursors (zero before the launch sequence)
};

// ---------------------------------------------------------------------------------------
// hand-rolled FP64 math: MUFU seed + polynomial refinement, no slow paths (arguments on this
// path are positive, finite and far from the denormal range; NaN still propagates).
// Measured on B200 (profiles/r01e_probe_pipes.json): both MUFU seeds are good to 2^-20.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    // one Halley step: y (1 + e/2 + 3 e^2 / 8), e = 1 - a y^2  (cubic convergence)
    double ay = a * y;
    double e = fma(-ay, y, 1.0);
    double p = fma(0.375, e, 0.5);
    double pe = p * e;
    return fma(y, pe, y);
}

// kMath: 0 = CUDA libm (sqrt, divide), 1 = MUFU seed + cubic-convergence step (errors ~1e-18),
// 2 = MUFU seed + one Newton step (quadratic convergence: relative error <= 3/8 e^2 = 3.1e-13 for
// the rsqrt, e^2 = 9.5e-13 for the reciprocal with the measured seed error e <= 2^-20; two and one
// FP64 instructions fewer).
template <int kMath>
__device__ __forceinline__ void radius(double u2, double rp, double &u, double &mur) {
    if (kMath == 1) {
        double y = fast_rsqrt(u2);
        u = u2 * y;
        mur = rp * y;
    } else if (kMath == 2) {
        double y0;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(u2));
        // u = a y0 (1 + h), mu_r = rp y0 (1 + h) with h = (1 - a y0^2) / 2 = 1/2 - (a y0)(y0 / 2);
        // y0 / 2 is an exponent decrement on the high word
        const double hy0 = __hiloint2double(__double2hiint(y0) - 0x00100000, __double2loint(y0));
        const double ay = u2 * y0;
        const double h = fma(-ay, hy0, 0.5);
        const double m0 = rp * y0;
        u = fma(ay, h, ay);
        mur = fma(m0, h, m0);
    } else {
        u = sqrt(u2);
        mur = rp / u;
    }
}

// 1/a: MUFU seed (rel. error e ~ 2^-20) then y (1 + e + e^2): cubic convergence
__device__ __forceinline__ double rcp_cubic(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double e = fma(-a, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

// 1/a: MUFU seed then y (1 + e): quadratic convergence (relative error e^2 <= 9.5e-13)
__device__ __forceinline__ double rcp_newton(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return fma(y, fma(-a, y, 1.0), y);
}

__device__ __forceinline__ double horner3(const double *c, double t) {
    return fma(fma(fma(c[3], t, c[2]), t, c[1]), t, c[0]);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// interval k with grid[k] <= b < grid[k+1], clamped to [0, n-2] (PCHIP extrapolates with its
// end polynomials: scipy PchipInterpolator(extrapolate=True), ccf_model.py:326, ccf_fit.py:193)
__device__ __forceinline__ int beta_interval(const double *grid, int n, double b) {
    int k = 0;
    for (int i = 1; i < n - 1; ++i) k += (b >= grid[i]) ? 1 : 0;
    return k;
}

__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
    double2 v;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int lds_s32(unsigned addr) {
    int v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// exp(r ln2/32) on |r| <= 1/2 -- coefficients in constant memory so that the FP64 pipe reads
// them as c[bank][offset] operands (a DFMA with a constant operand issues every 2.0 cycles,
// with three register operands every 2.2: profiles/r01e_probe_mix.txt).
//   kExpPoly  : degree-6 Taylor (truncation 3e-18)
//   kExpPoly5 : degree-5 economised fit at Chebyshev nodes (max relative error 1.4e-16; from a
//               60-digit mpmath solve), one FMA less
static __constant__ double kExpPoly[6] = {2.166084939249829e-02, 2.3459619820224677e-04, 1.6938509724371819e-06,
                                   9.172562701824643e-09, 3.9737099845494154e-11, 1.4345655584131932e-13};
static __constant__ double kExpPoly5[5] = {2.166084939249829e-02, 2.345961981994449e-04, 1.6938509724285119e-06,
                                    9.172607532092245e-09, 3.9737238568525983e-11};

// exp(-z2/2) from the 2^(j/32) table at 32-bit shared address `etab_s`
template <int kDeg>
__device__ __forceinline__ double gauss_tab(double z2, unsigned etab_s) {
    const double kMagic = 6755399441055744.0;   // 1.5 * 2^52
    const double kScale = -23.083120654223414;   // -16 log2(e)
    const double tn = fma(z2, kScale, kMagic);
    const int ni = __double2loint(tn);
    const double nf = tn - kMagic;
    const double r = fma(z2, kScale, -nf);
    double p;
    if (kDeg == 6) {
        p = fma(kExpPoly[5], r, kExpPoly[4]);
        p = fma(p, r, kExpPoly[3]);
        p = fma(p, r, kExpPoly[2]);
        p = fma(p, r, kExpPoly[1]);
        p = fma(p, r, kExpPoly[0]);
    } else {
        p = fma(kExpPoly5[4], r, kExpPoly5[3]);
        p = fma(p, r, kExpPoly5[2]);
        p = fma(p, r, kExpPoly5[1]);
        p = fma(p, r, kExpPoly5[0]);
    }
    p = fma(p, r, 1.0);
    const int n = max(ni >> 5, -1000);
    double t = lds_f64(etab_s + ((ni & (kExpTab - 1)) << 3));
    t = __hiloint2double(__double2hiint(t) + (n << 20), __double2loint(t));
    return p * t;
}

// the same with the argument pre-scaled: zs = z sqrt(16 log2 e), so that -z^2/2 = -(zs^2) ln2 / 32.
// zs^2 only ever appears inside an FMA (exact product, one instruction and four register reads
// fewer than forming z^2 first).  kGaussScale is folded into the sigma_v table by the caller.
constexpr double kGaussScale = 4.804489635145799;   // sqrt(16 log2(e))
static __constant__ float kExpPoly5f[5] = {2.166084939249829e-02f, 2.345961981994449e-04f, 1.6938509724285119e-06f,
                                    9.172607532092245e-09f, 3.9737238568525983e-11f};

// kDeg: 6 Taylor, 5 economised, 52 / 53 economised with the two / three highest Horner steps in FP32:
// the remainder satisfies |r ln2 / 32| <= 0.011, so the part of the polynomial multiplying r^3 (52) or
// r^2 (53) contributes at most 2e-7 / 6e-5 of the result and its FP32 rounding 1.2e-14 / 3.6e-12 -- the
// FP32 pipe and the two conversions (XU pipe) run beside the FP64 pipe, which is the bound.
template <int kDeg>
__device__ __forceinline__ double gauss_tab_scaled(double zs, unsigned etab_s) {
    const double kMagic = 6755399441055744.0;   // 1.5 * 2^52
    const double tn = fma(-zs, zs, kMagic);
    const int ni = __double2loint(tn);
    const double nf = tn - kMagic;
    const double r = fma(-zs, zs, -nf);
    double p;
    if (kDeg == 6) {
        p = fma(kExpPoly[5], r, kExpPoly[4]);
        p = fma(p, r, kExpPoly[3]);
        p = fma(p, r, kExpPoly[2]);
        p = fma(p, r, kExpPoly[1]);
        p = fma(p, r, kExpPoly[0]);
    } else if (kDeg == 52) {
        const float rf = __double2float_rn(r);
        const float pf = fmaf(fmaf(kExpPoly5f[4], rf, kExpPoly5f[3]), rf, kExpPoly5f[2]);
        p = fma((double)pf, r, kExpPoly5[1]);
        p = fma(p, r, kExpPoly5[0]);
    } else if (kDeg == 53) {
        const float rf = __double2float_rn(r);
        const float pf = fmaf(fmaf(fmaf(kExpPoly5f[4], rf, kExpPoly5f[3]), rf, kExpPoly5f[2]), rf, kExpPoly5f[1]);
        p = fma((double)pf, r, kExpPoly5[0]);
    } else {
        p = fma(kExpPoly5[4], r, kExpPoly5[3]);
        p = fma(p, r, kExpPoly5[2]);
        p = fma(p, r, kExpPoly5[1]);
        p = fma(p, r, kExpPoly5[0]);
    }
    p = fma(p, r, 1.0);
    const int n = max(ni >> 5, -1000);
    double t = lds_f64(etab_s + ((ni & (kExpTab - 1)) << 3));
    t = __hiloint2double(__double2hiint(t) + (n << 20), __double2loint(t));
    return p * t;
}

// exp(-z^2/2) from a 1024-entry table 2^(j/1024): zs = z sqrt(512 log2 e), -z^2/2 = -(zs^2) ln2 / 1024.
// The remainder |r ln2 / 1024| <= 3.4e-4 needs a degree-3 polynomial only (economised, max relative error
// 1.4e-16 -- the same as the degree-5 one on the 32-entry table): two FP64 instructions fewer.
//   kCvt: n = round(-zs^2) and its value back as a double through the conversion unit (F2I / I2F run on the
//   XU pipe beside the FP64 pipe; F2I saturates, so a huge z gives exp = 0 instead of a wrapped exponent)
//   instead of the 1.5 * 2^52 magic-number FMA + subtraction: one more FP64 instruction off the pipe.  zs^2 is
//   then rounded once (relative 1.1e-16, i.e. z^2/2 * 1.1e-16 relative in the result -- the conditioning of
//   exp(-z^2/2) itself, and what libm's exp(-0.5 * z * z) carries too).
constexpr double kGaussScaleBig = 27.178297609216609367;   // sqrt(512 log2(e))
static __constant__ double kExpPoly3[3] = {0.0006769015435155716, 2.2909785144706367e-07, 5.169222960550727e-11};

template <bool kCvt>
__device__ __forceinline__ double gauss_big(double zs, unsigned etab_s) {
    int ni;
    double r;
    if (kCvt) {
        const double q = zs * zs;
        asm("cvt.rni.s32.f64 %0, %1;" : "=r"(ni) : "d"(-q));
        r = -q - (double)ni;
    } else {
        const double kMagic = 6755399441055744.0;   // 1.5 * 2^52
        const double tn = fma(-zs, zs, kMagic);
        ni = __double2loint(tn);
        r = fma(-zs, zs, -(tn - kMagic));
    }
    double p = fma(kExpPoly3[2], r, kExpPoly3[1]);
    p = fma(p, r, kExpPoly3[0]);
    p = fma(p, r, 1.0);
    const int n = max(ni >> 10, -1000);
    double t = lds_f64(etab_s + ((ni & (kExpTabBig - 1)) << 3));
    t = __hiloint2double(__double2hiint(t) + (n << 20), __double2loint(t));
    return p * t;
}

// keep a value in a register: the compiler cannot re-derive the result of a volatile asm, so it
// stops re-materialising loop invariants (shared-window base, reciprocal spacing) inside the loop
__device__ __forceinline__ unsigned pin_u32(unsigned v) {
    unsigned r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ double pin_f64(double v) {
    double r;
    asm volatile("mov.f64 %0, %1;" : "=d"(r) : "d"(v));
    return r;
}

// ---------------------------------------------------------------------------------------
// pieces shared by both K1 kernels
// ---------------------------------------------------------------------------------------
// per-row scalars, scal[kNScal] in shared memory:
//   [0] f        template rescaling factor (ccf_model.py:606-613)
//   [1] aperp/f  [2] apar/f
//   [3] kappa    sigma_v iaH apar / f : displacement in u-units per unit x
//   [4] B        A_v / sigma_v, A_v = -growth / (3 iaH apar)               (:419, 435, 449)
//                (velocity template, growth_mode 2: A_v = fsigma8 / fs8t * growth_scale / apar, :439-443)
//   [5] G        iaH apar A_v / f   (dispersion / kaiser terms)
//   [6] apar     [7] aperp
//   [8] Av       empirical correction amplitude of the mean velocity (:453), times model bias / row bias
//                for the linear_bias matter model (delta carries 1 / bias, :367)
// warp 0 of the block computes the scalars of parameter row `pr` into shared `scal[kNScal]`
__device__ __forceinline__ void row_scalars_to_shared(const ModelDev &m, const double *pr, double *scal, int tid) {
    const double fs8 = pr[0], sigv = pr[2], aperp = pr[3], apar = pr[4], astar = pr[5];
    if (tid < 32) {
        const double eps = aperp / apar;
        double f;
        if (m.vel_indep_AP) {
            f = astar;
        } else {
            double part = 0.0;
            for (int i = tid; i < m.nresc; i += 32) {
                double mm = m.mu_resc[i];
                part += m.w_resc[i] * (apar * sqrt(1.0 + (1.0 - mm * mm) * (eps * eps - 1.0)));
            }
            f = warp_sum(part);
        }
        if (tid == 0) {
            const double iaHt = m.iaH * apar;
            // a bias given with the row (params.get('bias', model['bias']), :359, :430) rescales the
            // linear_bias profiles, which were tabulated with the model's bias; NaN = not given
            const double brow = pr[9];
            const bool has_b = m.lin_bias && (brow == brow);
            const double bs = has_b ? m.bias / brow : 1.0;
            double Av, G;
            if (m.growth_mode == 2) {
                Av = fs8 / m.fs8t * m.growth_scale / apar;                        // :439-443, :484
                G = iaHt * Av / f;
            } else {
                const double g = (m.growth_mode ? pr[1] * (has_b ? brow : m.bias) : fs8 / m.s8t) * bs;   // :425-435
                Av = -g / (3.0 * iaHt);
                G = -g / (3.0 * f);
            }
            scal[0] = f;
            scal[1] = aperp / f;
            scal[2] = apar / f;
            scal[3] = sigv * iaHt / f;
            scal[4] = Av / sigv;
            scal[5] = G;
            scal[6] = apar;
            scal[7] = aperp;
            scal[8] = pr[8] * bs;
            scal[9] = 0.0;
        }
    }
}

// xi(s_j, mu_k) staged in shared memory -> outputs: the xi block itself and / or its projection
// onto L <= 3 multipoles (one warp per s_j, lane-strided FMAs + shuffle reduction).
// ccf_model.py:824-825, utils.py:45-56, ccf_model.py:856-858.
__device__ __forceinline__ void write_outputs(const K1Args &a, const double *stage, long long row, int j0, int jn,
                                              int tid, int nthr, double *th = nullptr) {
    const int nmu = a.nmu;
    const int npairs = jn * nmu;
    if (a.xi_out) {
        double *xo = a.xi_out + (size_t)row * nmu * a.ns;
        for (int pidx = tid; pidx < npairs; pidx += nthr) {
            const int jl = pidx / nmu, k = pidx - jl * nmu;
            xo[(size_t)k * a.ns + j0 + jl] = stage[pidx];
        }
    }
    if (a.mult_out || th) {
        const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
        double *mo = a.mult_out ? a.mult_out + (size_t)row * a.L * a.ns : nullptr;
        for (int jl = warp; jl < jn; jl += nwarp) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (int k = lane; k < nmu; k += 32) {
                const double v = stage[jl * nmu + k];
                s0 = fma(a.wmu[k], v, s0);
                if (a.L > 1) s1 = fma(a.wmu[nmu + k], v, s1);
                if (a.L > 2) s2 = fma(a.wmu[2 * nmu + k], v, s2);
            }
            s0 = warp_sum(s0);
            if (a.L > 1) s1 = warp_sum(s1);
            if (a.L > 2) s2 = warp_sum(s2);
            if (lane == 0 && mo) {
                mo[j0 + jl] = s0;
                if (a.L > 1) mo[a.ns + j0 + jl] = s1;
                if (a.L > 2) mo[2 * a.ns + j0 + jl] = s2;
            }
            if (lane == 0 && th) {   // theory vector of this row, l-major (ccf_model.py:856-858)
                th[j0 + jl] = s0;
                if (a.L > 1) th[a.ns + j0 + jl] = s1;
                if (a.L > 2) th[2 * a.ns + j0 + jl] = s2;
            }
        }
    }
}

}  // namespace vb200
