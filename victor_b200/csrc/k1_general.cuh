// K1 (general): every model option of the path that the tuned streaming kernel does not cover --
//   rsd_model 'dispersion'                      victor/ccf_model.py:659-671
//   rsd_model 'kaiser' / 'euclid_special'       victor/ccf_model.py:692-741
//   anisotropic real-space input                victor/ccf_model.py:684-687  (assume_isotropic: False)
//   real-space ccf measured from data           victor/ccf_model.py:675-679  (realspace_ccf.from_data)
//   sigma_v(r, mu) dispersion templates         victor/ccf_model.py:654-655, 667-668 (3 template keys)
//   empirical correction of the mean velocity   victor/ccf_model.py:451-459  (with the dispersion / kaiser models;
//                                               the streaming case runs on the tuned kernel)
// and the streaming model combined with anisotropic, from-data or sigma_v(r, mu) input.  Same tiling as the
// tuned kernel (one block per parameter row x s-range, one thread per (s_j, mu_k) pair, velocity nodes in
// registers), the same hand-rolled rsqrt / reciprocal / exp (kFast) or CUDA libm (a test variant: both must
// pass parity), the same branch-free cell search; two velocity nodes in flight per thread.
#pragma once
#include "common.cuh"
#include "k2_chi2.cuh"

namespace vb200 {

// per-cell record of the general kernel:
//   xi_l cubics (3 x 4, unused ones zero) | V0 (4) | D0 (4) | SV (4) | origin | pad
// (V0, D0 times the row's velocity amplitude B, G or M G, whichever the model uses them with)
constexpr int kRecG = 26;
constexpr int kGU = 2;   // velocity nodes in flight per thread (general kernel)
template <int U>
struct NodeCount {
    static constexpr int value = U;
};
constexpr int kGXi = 0, kGV0 = 12, kGD0 = 16, kGSV = 20, kGOrg = 24;

__host__ __device__ inline size_t k1g_smem_bytes(int ncell, int jper, int nmu, int nbucket, int fitd = 0) {
    size_t d = (size_t)ncell * (kRecG + 1) + kExpTab + (size_t)jper * nmu + kNScal + fitd;
    return d * sizeof(double) + (size_t)nbucket * sizeof(int);
}

struct GenCtx {
    unsigned rec_s, bb_s;   // 32-bit shared-window addresses of the cell records and of the bucket table
    unsigned nbm1;
    const double *upper;
    int maxscan, has_flags, n_ell;
    int ell1, ell2;
    double inv_h;
    const double *sv2d, *sv_yb;
    int sv_ny;
};

// cubic c0 + c1 t + c2 t^2 + c3 t^3 whose coefficients sit at 32-bit shared address `addr` (two LDS.128)
__device__ __forceinline__ double cubic_s(unsigned addr, double t) {
    const double2 c01 = lds_f64x2(addr), c23 = lds_f64x2(addr + 16);
    return fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
}

// cell of coordinate u (shared address of its record) and the local coordinate inside it; below the
// first knot t = 0 (every spline is its boundary value there, FITPACK ext=3).  Same branch-free search as
// the tuned kernel: bucket = floor(u inv_h) through a round-down FMA onto 1.5 * 2^52 (u >= 0; NaN -> 0),
// one LDS.32 for the cell, a comparison scan only when some bucket holds a knot in its interior.
__device__ __forceinline__ unsigned locate(const GenCtx &g, double u, double &t) {
    const unsigned b = min((unsigned)__double2loint(__fma_rd(u, g.inv_h, 6755399441055744.0)), g.nbm1);
    int cell = lds_s32(g.bb_s + (b << 2));
    if (g.has_flags) {
        if (cell < 0) {
            cell &= ~kBucketFlag;
            for (int sc = 0; sc < g.maxscan; ++sc) cell += (u >= g.upper[cell]) ? 1 : 0;
        }
    }
    const unsigned ra = g.rec_s + cell * (kRecG * 8);
    const double tt = u - lds_f64(ra + kGOrg * 8);
    // t = max(t, 0) on the high word: a negative t becomes a denormal-sized positive number, 0 for the cubics
    t = __hiloint2double(max(__double2hiint(tt), 0), __double2loint(tt));
    return ra;
}

// normalised dispersion template at (cell record ra, local coordinate t, mu_r): the 1-D cubic, or
// for a sigma_v(r, mu) template the bicubic patch with mu clamped to the template's range
// (RectBivariateSpline.ev -> FITPACK bispeu clamps both arguments)
__device__ __forceinline__ double sv_at(const GenCtx &g, unsigned ra, double t, double mur) {
    if (g.sv_ny == 0) return cubic_s(ra + kGSV * 8, t);
    const int cell = (int)((ra - g.rec_s) / (kRecG * 8));
    const double *yb = g.sv_yb;
    double mc = mur;
    if (mc < yb[0]) mc = yb[0];
    if (mc > yb[g.sv_ny]) mc = yb[g.sv_ny];
    int yc = 0;
    for (int i = 1; i < g.sv_ny; ++i) yc += (mc >= yb[i]) ? 1 : 0;
    const double w = mc - yb[yc];
    const double *T = g.sv2d + ((size_t)cell * g.sv_ny + yc) * 16;
    double py[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) py[q] = horner3(T + 4 * q, w);
    return horner3(py, t);
}

__device__ __forceinline__ double legendre_even(int ell, double x) {
    // Horner forms of scipy.special.legendre(ell) (ccf_model.py:683, 687)
    const double x2 = x * x;
    if (ell == 0) return 1.0;
    if (ell == 2) return 1.5 * x2 - 0.5;
    return (4.375 * x2 - 3.75) * x2 + 0.375;
}

// real-space xi at (u, mu_r): sum_l xi_l(u) L_l(mu_r)                 ccf_model.py:681-687
__device__ __forceinline__ double xi_real(const GenCtx &g, unsigned ra, double t, double mur) {
    double xi = cubic_s(ra + kGXi * 8, t);
    if (g.n_ell > 1) xi += cubic_s(ra + (kGXi + 4) * 8, t) * legendre_even(g.ell1, mur);
    if (g.n_ell > 2) xi += cubic_s(ra + (kGXi + 8) * 8, t) * legendre_even(g.ell2, mur);
    return xi;
}

// arithmetic of the general kernel: hand-rolled (MUFU seed + cubic step, table exp) or libm
template <bool kFast>
struct GMath {
    // u = sqrt(u2), iu = 1 / u
    static __device__ __forceinline__ void root(double u2, double &u, double &iu) {
        if (kFast) {
            iu = fast_rsqrt(u2);
            u = u2 * iu;
        } else {
            u = sqrt(u2);
            iu = 1.0 / u;
        }
    }
    static __device__ __forceinline__ double div(double a, double b) { return kFast ? a * rcp_cubic(b) : a / b; }
    static __device__ __forceinline__ double gauss(double z2, unsigned etab_s) {
        return kFast ? gauss_tab<5>(z2, etab_s) : exp(-0.5 * z2);
    }
};

template <int kModel, bool kFast, bool kFuse>
__device__ __forceinline__ void general_body(const K1Args &a) {
    using M = GMath<kFast>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ModelDev &m = a.m;
    const int ncell = m.ncell, nx = m.nx;
    double *rec = reinterpret_cast<double *>(smem_raw);
    double *etab = rec + (size_t)ncell * kRecG;
    double *stage = etab + kExpTab;
    double *scal = stage + (size_t)a.jper * a.nmu;
    double *th = scal + kNScal;
    const int fitd = kFuse ? fused_fit_doubles(a.f.p) : 0;
    double *upper = th + fitd;
    int *bbase = reinterpret_cast<int *>(upper + ncell);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const long long row = blockIdx.x / a.nsplit;
    const int split = blockIdx.x - (int)(row * a.nsplit);
    const int j0 = split * a.jper;
    const int jn = min(a.jper, a.ns - j0);
    if (jn <= 0) return;

    const double *pr = a.params + row * kNPar;
    const double beta = m.beta_dependent ? pr[1] : m.beta_fixed;
    const double Mk = pr[6], Qk = pr[7];

    row_scalars_to_shared(m, pr, scal, tid);
    for (int i = tid; i < ncell; i += nthr) upper[i] = m.upper[i];
    for (int i = tid; i < m.nbucket; i += nthr) bbase[i] = m.bucket_base[i];
    if (tid < kExpTab) etab[tid] = m.exp_tab[tid];
    const unsigned etab_s = (unsigned)__cvta_generic_to_shared(etab);
    __syncthreads();   // the cell records below carry this row's velocity amplitude (and empirical correction)
    {
        const double Ae = m.v0b ? scal[8] : 0.0;
        // V0 and D0 are stored times the amplitude they are always used with: B = A_v / sigma_v in the mean of
        // the streaming pdf (:656), G = iaH A_v / f in the coordinate map and Jacobian of the dispersion model
        // (:660-671), M G in the kaiser / euclid_special forms (:698-714) -- one multiply less per evaluation
        const double vsc = kModel == kRsdStreaming ? scal[4] : (kModel == kRsdDispersion ? scal[5] : Mk * scal[5]);
        int kb = 0;
        double tb = 0.0;
        if (m.beta_dependent) {
            kb = beta_interval(m.beta_grid, m.nbeta, beta);
            tb = beta - m.beta_grid[kb];
        }
        const int per = ncell * 4;
        const size_t ell_stride = (size_t)(m.nbeta - 1) * 4 * per;
        for (int i = tid; i < per; i += nthr) {
            const int cell = i >> 2, c = i & 3;
            double *r = rec + cell * kRecG;
            for (int l = 0; l < kMaxPoles; ++l) {
                double v = 0.0;
                if (l < m.n_ell) {
                    const double *tab = m.xi_tab + l * ell_stride + (size_t)kb * 4 * per;
                    v = fma(fma(fma(tab[3 * per + i], tb, tab[2 * per + i]), tb, tab[per + i]), tb, tab[i]);
                }
                r[kGXi + 4 * l + c] = v;
            }
            if (m.vd_beta_dep) {   // linear_bias: V0, D0 follow the monopole's beta dependence
                const double *tv = m.v0 + (size_t)kb * 4 * per, *td = m.d0 + (size_t)kb * 4 * per;
                r[kGV0 + c] = vsc * fma(fma(fma(tv[3 * per + i], tb, tv[2 * per + i]), tb, tv[per + i]), tb, tv[i]);
                r[kGD0 + c] = vsc * fma(fma(fma(td[3 * per + i], tb, td[2 * per + i]), tb, td[per + i]), tb, td[i]);
            } else if (m.v0b) {    // empirical correction (1 + Av delta(r)) of the mean velocity, ccf_model.py:451-459
                r[kGV0 + c] = vsc * fma(Ae, m.v0b[i], m.v0[i]);
                r[kGD0 + c] = vsc * fma(Ae, m.d0b[i], m.d0[i]);
            } else {
                r[kGV0 + c] = vsc * m.v0[i];
                r[kGD0 + c] = vsc * m.d0[i];
            }
            r[kGSV + c] = m.sv[i];
            if (c == 0) {
                r[kGOrg] = m.origin[cell];
                r[kGOrg + 1] = 0.0;
            }
        }
    }
    __syncthreads();

    const double f = scal[0], sperp_f = scal[1], spar_f = scal[2], kappa = scal[3];
    const double apar = scal[6];
    GenCtx g;
    g.rec_s = (unsigned)__cvta_generic_to_shared(rec);
    g.bb_s = (unsigned)__cvta_generic_to_shared(bbase);
    g.nbm1 = (unsigned)(m.nbucket - 1);
    g.upper = upper;
    g.maxscan = m.maxscan;
    g.has_flags = a.has_flags;
    g.n_ell = m.n_ell;
    g.ell1 = m.ells[1];
    g.ell2 = m.ells[2];
    g.inv_h = m.inv_h;
    g.sv2d = m.sv2d;
    g.sv_yb = m.sv_yb;
    g.sv_ny = m.sv_ny;
    const int nmu = a.nmu;
    const int npairs = jn * nmu;
    const double f_over_apar = f / apar;

    for (int pidx = tid; pidx < npairs; pidx += nthr) {
        const int jl = pidx / nmu, k = pidx - jl * nmu;
        const double sj = a.s[j0 + jl];
        const int km = a.pairwise ? j0 + jl : k;
        const double Sperp = sj * a.sqmu[km] * sperp_f;
        const double Spar = sj * a.mu[km] * spar_f;
        const double Sperp2 = Sperp * Sperp;
        const double rt_data = sj * a.sqmu[km];   // s_perp / aperp in real units (from_data, :676)
        double result;

        // real-space xi for a point with (u-unit) line-of-sight separation rp
        auto xi_at = [&](double rp, double u, double mur, unsigned rcell, double t) {
            if (!m.from_data) return xi_real(g, rcell, t, mur);
            const double rpd = rp * f_over_apar;                       // r_par / apar      (:675)
            double rd, ird;
            M::root(rpd * rpd + rt_data * rt_data, rd, ird);           // (:677)
            double td;
            const unsigned rc = locate(g, rd, td);
            return xi_real(g, rc, td, kFast ? rpd * ird : rpd / rd);   // (:678-687)
        };

        // U consecutive velocity nodes of this (s_j, mu_k) pair, stage by stage, so that U independent
        // dependency chains are in flight per warp (the fixed-point iteration of the dispersion model is one
        // long chain per node: rsqrt -> cell -> cubic -> reciprocal, six times over)
        auto stream_nodes = [&](auto Uc, int mi, double acc) {
            constexpr int U = decltype(Uc)::value;
            double xm[U], rp[U], u[U], iu[U], mur[U], t[U], isv[U], sv[U], z[U], xi[U];
            unsigned rc[U];
#pragma unroll
            for (int i = 0; i < U; ++i) {
                xm[i] = a.xw[mi + i];
                rp[i] = Spar - xm[i] * kappa;                          // :648
                M::root(Sperp2 + rp[i] * rp[i], u[i], iu[i]);         // :651
                mur[i] = kFast ? rp[i] * iu[i] : rp[i] / u[i];         // :652
            }
#pragma unroll
            for (int i = 0; i < U; ++i) rc[i] = locate(g, u[i], t[i]);
#pragma unroll
            for (int i = 0; i < U; ++i) {
                sv[i] = sv_at(g, rc[i], t[i], mur[i]);                 // :654-655
                isv[i] = kFast ? rcp_cubic(sv[i]) : 1.0 / sv[i];
            }
#pragma unroll
            for (int i = 0; i < U; ++i) {
                const double d = fma(-cubic_s(rc[i] + kGV0 * 8, t[i]), mur[i], xm[i]);   // x - (B V0)(u) mu_r
                z[i] = kFast ? d * isv[i] : d / sv[i];                 // :656
            }
#pragma unroll
            for (int i = 0; i < U; ++i) xi[i] = xi_at(rp[i], u[i], mur[i], rc[i], t[i]);
#pragma unroll
            for (int i = 0; i < U; ++i) {
                const double wm = a.xw[kMaxNx + mi + i];
                const double pdf = M::gauss(z[i] * z[i], etab_s);
                acc += kFast ? wm * (1.0 + xi[i]) * pdf * isv[i] : wm * (1.0 + xi[i]) * pdf / sv[i];   // :690
            }
            return acc;
        };

        auto disp_nodes = [&](auto Uc, int mi, double acc, double first, double ifirst) {
            constexpr int U = decltype(Uc)::value;
            double xm[U], num[U], rp[U], u[U], iu[U], t[U];
            unsigned rc[U];
#pragma unroll
            for (int i = 0; i < U; ++i) {
                xm[i] = a.xw[mi + i];
                num[i] = Spar - xm[i] * kappa;
                rp[i] = kFast ? num[i] * ifirst : num[i] / first;
            }
            for (int it = 0; it < m.niter; ++it) {
#pragma unroll
                for (int i = 0; i < U; ++i) M::root(Sperp2 + rp[i] * rp[i], u[i], iu[i]);
#pragma unroll
                for (int i = 0; i < U; ++i) rc[i] = locate(g, u[i], t[i]);
#pragma unroll
                for (int i = 0; i < U; ++i)
                    rp[i] = kFast ? num[i] * rcp_cubic(fma(cubic_s(rc[i] + kGV0 * 8, t[i]), iu[i], 1.0))
                                  : num[i] / (1.0 + cubic_s(rc[i] + kGV0 * 8, t[i]) / u[i]);
            }
            double mur[U];
#pragma unroll
            for (int i = 0; i < U; ++i) {
                M::root(Sperp2 + rp[i] * rp[i], u[i], iu[i]);
                mur[i] = kFast ? rp[i] * iu[i] : rp[i] / u[i];
            }
#pragma unroll
            for (int i = 0; i < U; ++i) rc[i] = locate(g, u[i], t[i]);
#pragma unroll
            for (int i = 0; i < U; ++i) {
                const double sv = sv_at(g, rc[i], t[i], mur[i]);       // :667-668
                const double v0 = cubic_s(rc[i] + kGV0 * 8, t[i]);
                const double v0u = kFast ? v0 * iu[i] : v0 / u[i];
                const double jd = fma(mur[i] * mur[i], cubic_s(rc[i] + kGD0 * 8, t[i]) - v0u, 1.0 + v0u);
                const double xi = xi_at(rp[i], u[i], mur[i], rc[i], t[i]);
                const double wm = a.xw[kMaxNx + mi + i];
                if (kFast) {
                    const double isv = rcp_cubic(sv);
                    const double z = xm[i] * isv;
                    acc += wm * (1.0 + xi) * rcp_cubic(jd) * M::gauss(z * z, etab_s) * isv;
                } else {
                    const double z = xm[i] / sv;
                    acc += wm * (1.0 + xi) * (1.0 / jd) * exp(-0.5 * z * z) / sv;
                }
            }
            return acc;
        };

        if (kModel == kRsdStreaming) {
            double acc = 0.0;
            int mi = 0;
            for (; mi + kGU <= nx; mi += kGU) acc = stream_nodes(NodeCount<kGU>{}, mi, acc);
            for (; mi < nx; ++mi) acc = stream_nodes(NodeCount<1>{}, mi, acc);
            result = acc - 1.0;
        } else if (kModel == kRsdDispersion) {
            // ccf_model.py:659-671
            double Strue, iS;
            M::root(Sperp2 + Spar * Spar, Strue, iS);
            double t0;
            const unsigned r0 = locate(g, Strue, t0);
            const double first = kFast ? fma(cubic_s(r0 + kGV0 * 8, t0), iS, 1.0)
                                       : 1.0 + cubic_s(r0 + kGV0 * 8, t0) / Strue;
            const double ifirst = kFast ? rcp_cubic(first) : 0.0;
            double acc = 0.0;
            int mi = 0;
            for (; mi + kGU <= nx; mi += kGU) acc = disp_nodes(NodeCount<kGU>{}, mi, acc, first, ifirst);
            for (; mi < nx; ++mi) acc = disp_nodes(NodeCount<1>{}, mi, acc, first, ifirst);
            result = acc - 1.0;
        } else {
            // kaiser / euclid_special: ccf_model.py:692-741
            double rp = Spar;
            double u, t;
            unsigned rc;
            double iu;
            if (m.kaiser_shift) {
                double Strue;
                M::root(Sperp2 + Spar * Spar, Strue, iu);
                rc = locate(g, Strue, t);
                rp = kFast ? Spar * rcp_cubic(fma(cubic_s(rc + kGV0 * 8, t), iu, 1.0))
                           : Spar / (1.0 + cubic_s(rc + kGV0 * 8, t) / Strue);
                for (int it = 0; it < m.niter; ++it) {
                    M::root(Sperp2 + rp * rp, u, iu);
                    rc = locate(g, u, t);
                    rp = kFast ? Spar * rcp_cubic(fma(cubic_s(rc + kGV0 * 8, t), iu, 1.0))
                               : Spar / (1.0 + cubic_s(rc + kGV0 * 8, t) / u);
                }
            }
            M::root(Sperp2 + rp * rp, u, iu);
            const double mur = kFast ? rp * iu : rp / u;
            rc = locate(g, u, t);
            const double v0u = kFast ? cubic_s(rc + kGV0 * 8, t) * iu : cubic_s(rc + kGV0 * 8, t) / u;
            const double ca = (m.rsd_model == kRsdEuclid) ? 3.0 : 1.0, cb = (m.rsd_model == kRsdEuclid) ? 2.0 : 1.0;
            const double J = ca * v0u + cb * Qk * mur * mur * (cubic_s(rc + kGD0 * 8, t) - v0u);
            const double xi = xi_at(rp, u, mur, rc, t);
            if (m.rsd_model == kRsdEuclid || m.kaiser_approx)
                result = Mk * xi - J;
            else
                result = M::div(1.0 + Mk * xi, 1.0 + J) - 1.0;
        }
        stage[pidx] = result;
    }
    __syncthreads();
    write_outputs(a, stage, row, j0, jn, tid, nthr, kFuse ? th : nullptr);
    if (kFuse) {   // one block per row: finish with chi2 and lnL (ccf_fit.py:349-354, 441-483)
        __syncthreads();
        block_chi2(a.f, pr[1], th, th + ((a.f.p + 1) & ~1), row, a.chi2, a.lnl, tid, nthr);
    }
}

// Velocity-integral models: 64 registers, four blocks per SM (measured best: dispersion 547k evals/s against
// 535k at the 80 registers ptxas picks unconstrained).  The kaiser forms have 3000 points per row and no
// velocity loop; they run best with ptxas's own choice (62 registers), so they get their own entry point.
template <int kModel, bool kFast, bool kFuse = false>
__global__ void __launch_bounds__(256, 4) k_multipoles_general(const __grid_constant__ K1Args a) {
    general_body<kModel, kFast, kFuse>(a);
}

template <bool kFast, bool kFuse = false>
__global__ void __launch_bounds__(256) k_multipoles_kaiser(const __grid_constant__ K1Args a) {
    general_body<kRsdKaiser, kFast, kFuse>(a);
}

}  // namespace vb200
