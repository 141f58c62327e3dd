// K1 (tuned): xi(s, mu) and its multipoles for the velocity-integral models on model coordinates with an
// isotropic sigma_v(r) template -- the BOSS DR12 CMASS setups:
//   Gaussian streaming model   victor/ccf_model.py:646-658   (isotropic real-space xi: BASELINE configs 1-3, 5;
//                                                              or xi_0 + xi_2 L_2 (+ xi_4 L_4), :684-687)
//   dispersion model           victor/ccf_model.py:659-671   (fixed-point coordinate map + Jacobian)
//
// One thread block per (parameter row, s-bin range).  Replaces
//   CCFModel.theory_xi                    victor/ccf_model.py:589-671, 681-690
//   CCFModel.theory_multipoles            victor/ccf_model.py:816-825 + victor/utils.py:45-56
//   CCFModel.theory_multipole_vector      victor/ccf_model.py:856-858
// The block first turns the host tables into its row's own cell table in shared memory
// (beta-Horner of the xi^r power table, velocity amplitude folded into V0), then every thread
// owns (s_j, mu_k) pairs and runs the velocity quadrature over x_m in registers; xi(s_j, mu_k)
// is staged in shared memory and projected onto the multipoles with warp shuffles.
#pragma once
#include "common.cuh"
#include "k2_chi2.cuh"

#ifndef VB200_TAIL2
#define VB200_TAIL2 1
#endif

namespace vb200 {

// Shared memory (dynamic), see k1_smem_bytes():
//   rec[ncell][kRec]  per-row cell records (16 B aligned), doubles
//                       [0..3]   xi_0 + 1
//                       [4..7]   streaming: B V0      dispersion: G V0          (B = A_v / sigma_v, G = iaH' A_v / f)
//                       [8..11]  SV (kFast: divided by sqrt(16 log2 e))
//                       [12]     origin   [13] pad
//                       dispersion only: [14..17] G D0 + 1
//                       then xi_2, xi_4 (anisotropic real-space input), 4 each
//                     streaming + isotropic: stride 112 B = 28 banks, so 8 consecutive cells tile the 32 banks
//   etab[32 | 1024] 2^(j/32) or 2^(j/1024) (C::kTab)
//   stage[jper*nmu] xi(s_j, mu_k) of this block
//   scal[kNScal]    per-row scalars
//   th[fitd]        fused likelihood epilogue only: theory / residual vector and per-warp partial sums
//   upper[ncell]    upper knot of each cell (slow path of the cell search only)
//   int bbase[nbucket]  bucket -> first cell; bit 31 set when a knot lies strictly inside the bucket
constexpr int kRec = 14;

__host__ __device__ inline int k1_rec_doubles(int rsd_model, int n_ell) {
    return kRec + (rsd_model == kRsdDispersion ? 4 : 0) + 4 * (n_ell - 1);
}

__host__ __device__ inline size_t k1_smem_bytes(int ncell, int jper, int nmu, int nbucket, int fitd = 0, int rec = kRec,
                                                int tab = kExpTab) {
    size_t d = (size_t)ncell * (rec + 1) + tab + (size_t)jper * nmu + kNScal + fitd;
    return d * sizeof(double) + (size_t)nbucket * sizeof(int);
}

// Compile-time variant.
//   kFast  : hand-rolled rsqrt / rcp / exp (else CUDA libm).
//   kNewton: 3 = cubic-convergence refinement of the MUFU seeds, 2 = one Newton step.
//   kFlags : some bucket holds a knot in its interior, so the cell search may need the
//            comparison path (non-lattice knot sets).
//   kU     : velocity nodes processed together per loop trip (instruction-level parallelism; a warp
//            must keep >= 4 independent DFMAs in flight to reach the FP64 issue rate).
//   kExp   : exp remainder polynomial: 5 = economised degree 5 on the 32-entry table (default), 6 = Taylor (libm-free
//            reference form), 3 = degree 3 on a 1024-entry table with the range reduction through the conversion unit
//            (gauss_big; measured no faster, kept as the second form).  The FP32-tail forms (52 / 53) and the
//            magic-number big-table form (30) measured in round 2 remain as device functions for the probes library.
//   kModel : kRsdStreaming or kRsdDispersion.
//   kNEll  : real-space multipoles in xi(r, mu_r): 1 (isotropic), 2 (0, 2) or 3 (0, 2, 4).
//   kMinBlocks : 4 -> 64 registers per thread, 3 -> 80.
//   kFromData : model['realspace_ccf']['from_data'] (ccf_model.py:675-679): the real-space ccf was measured in the
//            fiducial cosmology, so xi is looked up at (r_par / apar, s_perp / aperp) -- a second radius and cell per node.
template <bool kFast_, bool kFlags_, int kU_, int kExp_, int kNewton_ = 3, int kModel_ = kRsdStreaming, int kNEll_ = 1,
          int kMinBlocks_ = 4, bool kFromData_ = false>
struct K1Cfg {
    static constexpr bool kFromData = kFromData_;    // real-space ccf measured from data: xi at fiducial coordinates
    static constexpr int kMinBlocks = kMinBlocks_;   // resident blocks per SM the register budget is set for
    static constexpr bool kFast = kFast_, kFlags = kFlags_;
    static constexpr int kU = kU_, kExp = kExp_;
    static constexpr int kMath = kFast_ ? (kNewton_ == 2 ? 2 : 1) : 0;
    static constexpr int kModel = kModel_, kNEll = kNEll_;
    static constexpr bool kBigTab = kExp_ == 3 || kExp_ == 30;
    static constexpr int kTab = kBigTab ? kExpTabBig : kExpTab;
    static constexpr double kScale = kBigTab ? kGaussScaleBig : kGaussScale;   // folded into the sigma_v table
    static __device__ __forceinline__ double gauss(double zs, unsigned etab_s) {
        if (kExp_ == 3) return gauss_big<true>(zs, etab_s);
        if (kExp_ == 30) return gauss_big<false>(zs, etab_s);
        return gauss_tab_scaled<kExp_>(zs, etab_s);
    }
    static constexpr bool kDisp = kModel_ == kRsdDispersion;
    static constexpr int kD1 = 14;                       // dispersion: G D0 + 1
    static constexpr int kXiHi = kDisp ? 18 : 14;        // xi_2 (then xi_4)
    static constexpr int kRecD = kXiHi + 4 * (kNEll_ - 1);
};

// per-thread loop invariants of the quadrature
struct QuadCtx {
    double kappa, Spar, Sperp2, inv_h;
    unsigned nbm1, bb_s, rec_s, etab_s;
    const double *upper;
    int maxscan;
    double first, ifirst;   // dispersion: 1 + G V0(S) / S at the redshift-space point itself, and its reciprocal
    int niter;
    double f_over_apar, rt2;   // kFromData: u-units -> fiducial Mpc/h along the line of sight; (s_perp / aperp)^2
};

// shared address of the cell record that holds coordinate u
template <class C>
__device__ __forceinline__ unsigned cell_record(const QuadCtx &q, double u) {
    // bucket = floor(u * inv_h) through a round-down FMA onto 1.5 * 2^52 (u >= 0; NaN -> 0)
    const unsigned b = min((unsigned)__double2loint(__fma_rd(u, q.inv_h, 6755399441055744.0)), q.nbm1);
    int cell = lds_s32(q.bb_s + (b << 2));
    if (C::kFlags) {
        if (cell < 0) {  // a knot lies inside this bucket: finish the search by comparison
            cell &= ~kBucketFlag;
            for (int sc = 0; sc < q.maxscan; ++sc) cell += (u >= q.upper[cell]) ? 1 : 0;
        }
    }
    return q.rec_s + cell * (C::kRecD * 8);
}

// local coordinate in the cell: t = max(u - origin, 0).  Below the first knot every spline is its boundary
// value (ext=3), which is the first cell's cubic at t = 0.  Done on the high word with an integer max: a
// negative t becomes a positive denormal-sized number, i.e. 0 for the cubic.
__device__ __forceinline__ double cell_coord(unsigned ra, double u) {
    const double tt = u - lds_f64(ra + 96);
    return __hiloint2double(max(__double2hiint(tt), 0), __double2loint(tt));
}

__device__ __forceinline__ double cubic_at(unsigned addr, double t) {
#ifdef VB200_DIAG_LDS   // diagnostic build only (wrong results): half the coefficient bytes, same arithmetic
    const double2 c01 = lds_f64x2(addr), c23 = make_double2(c01.y, c01.x);
#else
    const double2 c01 = lds_f64x2(addr), c23 = lds_f64x2(addr + 16);
#endif
    return fma(fma(fma(c23.y, t, c23.x), t, c01.y), t, c01.x);
}

// 1 + xi^r(u, mu_r) = (xi_0 + 1)(u) + xi_2(u) L_2(mu_r) + xi_4(u) L_4(mu_r)        ccf_model.py:681-687
template <class C>
__device__ __forceinline__ double xi_plus_one(unsigned ra, double t, double mur) {
    double xi1 = cubic_at(ra, t);
    if (C::kNEll > 1) {
        const double x2 = mur * mur;
        xi1 = fma(cubic_at(ra + C::kXiHi * 8, t), fma(1.5, x2, -0.5), xi1);
        if (C::kNEll > 2) xi1 = fma(cubic_at(ra + (C::kXiHi + 4) * 8, t), fma(fma(4.375, x2, -3.75), x2, 0.375), xi1);
    }
    return xi1;
}

// 1 + xi^r for a node at line-of-sight separation rp (u-units): at its own record / coordinate / mu_r, or, for a real-space
// ccf measured from data, at the fiducial-cosmology point (r_par / apar, s_perp / aperp)            ccf_model.py:675-687
template <class C>
__device__ __forceinline__ double xi_node(const QuadCtx &q, unsigned ra, double t, double mur, double rp) {
    if (!C::kFromData) return xi_plus_one<C>(ra, t, mur);
    const double rpd = rp * q.f_over_apar;                 // :675
    const double rd2 = fma(rpd, rpd, q.rt2);               // :677
    double rd, mud;
    radius<C::kMath>(rd2, rpd, rd, mud);                   // :678
    const unsigned rc = cell_record<C>(q, rd);
    return xi_plus_one<C>(rc, cell_coord(rc, rd), mud);
}

// U consecutive velocity nodes of one (s_j, mu_k) pair, written stage by stage so that the U
// dependency chains are interleaved in program order (DFMA latency 8.5 cycles, issue every 2.2).
template <class C, int U>
__device__ __forceinline__ double quad_nodes(const K1Args &a, const QuadCtx &q, int mi, double acc) {
    double xm[U], u[U], mur[U], t[U], rq[U], z2[U], g[U];   // kFast: z2 holds zs = z sqrt(16 log2 e), not z^2
    double rpk[C::kFromData ? U : 1];
    unsigned ra[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
        xm[i] = a.xw[mi + i];                                 // uniform: constant-bank read
        const double rp = fma(-xm[i], q.kappa, q.Spar);       // ccf_model.py:648-650
        const double u2 = fma(rp, rp, q.Sperp2);              // :651
        radius<C::kMath>(u2, rp, u[i], mur[i]);               // :651-652
        if (C::kFromData) rpk[i] = rp;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) ra[i] = cell_record<C>(q, u[i]);
#pragma unroll
    for (int i = 0; i < U; ++i) t[i] = cell_coord(ra[i], u[i]);
#pragma unroll
    for (int i = 0; i < U; ++i) {
        const double sv = cubic_at(ra[i] + 64, t[i]);   // :654-655
        rq[i] = C::kMath == 1 ? rcp_cubic(sv) : (C::kMath == 2 ? rcp_newton(sv) : 1.0 / sv);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
        const double vb = cubic_at(ra[i] + 32, t[i]);   // :635, :656
        const double z = fma(-vb, mur[i], xm[i]) * rq[i];   // kFast: rq = sqrt(16 log2 e) / SV (scaled table)
        z2[i] = C::kFast ? z : z * z;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) g[i] = C::kFast ? C::gauss(z2[i], q.etab_s) : exp(-0.5 * z2[i]);
#pragma unroll
    for (int i = 0; i < U; ++i) {
        const double xi1 = xi_node<C>(q, ra[i], t[i], mur[i], C::kFromData ? rpk[i] : 0.0);   // :621, :675-687
        acc = fma(a.xw[kMaxNx + mi + i] * (xi1 * rq[i]), g[i], acc);                      // :690
    }
    return acc;
}

// The same for the dispersion model (ccf_model.py:659-671): per node the real-space line-of-sight
// separation solves r_par = (s_par - x sigma_v iaH') / (1 + iaH' v_r(r) / r) by `niter` fixed-point
// iterations from the value at the redshift-space point, then the integrand carries the Jacobian
// 1 + v/r + mu_r^2 (v' - v/r) and a Gaussian in x / SV(r, mu_r).
// kFast: the iteration is written rp <- num u / (u + G V0(u)) (no 1/u).
template <class C, int U>
__device__ __forceinline__ double disp_nodes(const K1Args &a, const QuadCtx &q, int mi, double acc) {
    double xm[U], num[U], rp[U], u[U], t[U];
    unsigned ra[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
        xm[i] = a.xw[mi + i];
        num[i] = fma(-xm[i], q.kappa, q.Spar);
        rp[i] = C::kFast ? num[i] * q.ifirst : num[i] / q.first;   // :660-661
    }
    // :662-664.  The map is NOT a contraction everywhere in the prior box: at large separations and high growth
    // rates |d rp_new / d rp| exceeds 1 (up to ~25 per iteration among the rows of the bench batch), so a rounding
    // error made in an early iteration is amplified by up to 1e6 by the fifth.  Every iteration therefore refines
    // its MUFU seeds to full precision (C::kMath == 1: cubic steps); one-Newton-step seeds (1e-12) end at 5e-6
    // relative in the worst rows (profiles/r02c_parity_report.jsonl) and exist as a measurement variant only.
    for (int it = 0; it < q.niter; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const double u2 = fma(rp[i], rp[i], q.Sperp2);
            if (C::kMath == 1) {
                double y0;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(u2));
                const double ay = u2 * y0;
                const double e = fma(-ay, y0, 1.0);
                const double pe = fma(0.375, e, 0.5) * e;      // Halley: u = a y0 (1 + e/2 + 3 e^2/8)
                u[i] = fma(ay, pe, ay);
            } else if (C::kMath == 2) {
                double y0;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(u2));
                const double hy0 = __hiloint2double(__double2hiint(y0) - 0x00100000, __double2loint(y0));
                const double ay = u2 * y0;
                const double h = fma(-ay, hy0, 0.5);
                u[i] = fma(ay, h, ay);
            } else {
                u[i] = sqrt(u2);
            }
        }
#pragma unroll
        for (int i = 0; i < U; ++i) ra[i] = cell_record<C>(q, u[i]);
#pragma unroll
        for (int i = 0; i < U; ++i) t[i] = cell_coord(ra[i], u[i]);
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const double gv = cubic_at(ra[i] + 32, t[i]);
            if (C::kFast) {
                const double den = gv + u[i];
                double y;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
                const double e = fma(-den, y, 1.0);
                const double p = (num[i] * u[i]) * y;
                rp[i] = fma(p, C::kMath == 1 ? fma(e, e, e) : e, p);
            } else {
                rp[i] = num[i] / (1.0 + gv / u[i]);
            }
        }
    }
    double y[U], mur[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {                                  // :665-666
        const double u2 = fma(rp[i], rp[i], q.Sperp2);
        if (C::kMath == 1) {
            y[i] = fast_rsqrt(u2);
            u[i] = u2 * y[i];
            mur[i] = rp[i] * y[i];
        } else if (C::kMath == 2) {
            double y0;
            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(u2));
            const double hy0 = __hiloint2double(__double2hiint(y0) - 0x00100000, __double2loint(y0));
            const double ay = u2 * y0;
            const double h = fma(-ay, hy0, 0.5);
            const double m0 = rp[i] * y0;
            u[i] = fma(ay, h, ay);
            mur[i] = fma(m0, h, m0);
            y[i] = fma(y0, h, y0);
        } else {
            u[i] = sqrt(u2);
            mur[i] = rp[i] / u[i];
            y[i] = 1.0 / u[i];
        }
    }
#pragma unroll
    for (int i = 0; i < U; ++i) ra[i] = cell_record<C>(q, u[i]);
#pragma unroll
    for (int i = 0; i < U; ++i) t[i] = cell_coord(ra[i], u[i]);
    double rq[U], rj[U], g[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
        const double sv = cubic_at(ra[i] + 64, t[i]);             // :667-668
        rq[i] = C::kMath == 1 ? rcp_cubic(sv) : (C::kMath == 2 ? rcp_newton(sv) : 1.0 / sv);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {                                  // Jacobian, :669-670
        const double gv = cubic_at(ra[i] + 32, t[i]);
        const double w1 = C::kFast ? fma(gv, y[i], 1.0) : 1.0 + gv / u[i];   // 1 + v / r
        const double d1 = cubic_at(ra[i] + C::kD1 * 8, t[i]);                // 1 + v'
        const double jd = fma(mur[i] * mur[i], d1 - w1, w1);
        rj[i] = C::kMath == 1 ? rcp_cubic(jd) : (C::kMath == 2 ? rcp_newton(jd) : 1.0 / jd);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
        const double z = xm[i] * rq[i];                            // kFast: rq = sqrt(16 log2 e) / SV
        g[i] = C::kFast ? C::gauss(z, q.etab_s) : exp(-0.5 * z * z);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
        const double xi1 = xi_node<C>(q, ra[i], t[i], mur[i], rp[i]);
        acc = fma(a.xw[kMaxNx + mi + i] * (xi1 * rq[i]) * rj[i], g[i], acc);   // :671, :690
    }
    return acc;
}

template <class C, int U>
__device__ __forceinline__ double nodes(const K1Args &a, const QuadCtx &q, int mi, double acc) {
    return C::kDisp ? disp_nodes<C, U>(a, q, mi, acc) : quad_nodes<C, U>(a, q, mi, acc);
}

// this row's cell records in shared memory: xi^r(u; beta) from the beta power table (+1 folded into c0),
// amplitude * V0(u), SV(u), origin (dispersion: G D0 + 1 too); `scal` = row_scalars_to_shared's output
// kRaw: V0 (and D0) are stored WITHOUT the row's velocity amplitude, so that the build need not wait for the row
// scalars (k_small overlaps the two); scale_cell_records() applies the amplitude afterwards -- the same products.
template <class C, bool kRaw = false>
__device__ __forceinline__ void build_cell_records(const ModelDev &m, const double *scal, double beta, double *rec,
                                                   int tid, int nthr) {
    constexpr int kR = C::kRecD;
    const int ncell = m.ncell;
    const double amp = kRaw ? 1.0 : (C::kDisp ? scal[5] : scal[4]);
    int kb = 0;
    double tb = 0.0;
    if (m.beta_dependent) {
        kb = beta_interval(m.beta_grid, m.nbeta, beta);
        tb = beta - m.beta_grid[kb];
    }
    const int per = ncell * 4;
    const size_t ell_stride = (size_t)(m.nbeta - 1) * 4 * per;
    const double *tab = m.xi_tab + (size_t)kb * 4 * per;  // [q][cell][4], ell index 0
    for (int i = tid; i < per; i += nthr) {
        const int cell = i >> 2, c = i & 3;
        double v = fma(fma(fma(tab[3 * per + i], tb, tab[2 * per + i]), tb, tab[per + i]), tb, tab[i]);
        if (c == 0) v += 1.0;
        double *r = rec + cell * kR;
        r[c] = v;
#pragma unroll
        for (int l = 1; l < C::kNEll; ++l) {
            const double *tl = tab + l * ell_stride;
            r[C::kXiHi + 4 * (l - 1) + c] =
                fma(fma(fma(tl[3 * per + i], tb, tl[2 * per + i]), tb, tl[per + i]), tb, tl[i]);
        }
        double v0 = m.v0[i];
        double d0 = C::kDisp ? m.d0[i] : 0.0;
        if (m.vd_beta_dep) {   // linear_bias matter model: V0 follows the monopole's beta dependence (ccf_model.py:358-370)
            const double *tv = m.v0 + (size_t)kb * 4 * per;
            v0 = fma(fma(fma(tv[3 * per + i], tb, tv[2 * per + i]), tb, tv[per + i]), tb, tv[i]);
            if (C::kDisp) {
                const double *td = m.d0 + (size_t)kb * 4 * per;
                d0 = fma(fma(fma(td[3 * per + i], tb, td[2 * per + i]), tb, td[per + i]), tb, td[i]);
            }
        } else if (m.v0b) {    // empirical correction (1 + Av delta(r)) of the mean velocity (:451-455)
            v0 = fma(scal[8], m.v0b[i], v0);
            if (C::kDisp) d0 = fma(scal[8], m.d0b[i], d0);
        }
        r[4 + c] = kRaw ? v0 : amp * v0;
        if (C::kDisp) r[C::kD1 + c] = kRaw ? d0 : ((c == 0) ? fma(amp, d0, 1.0) : amp * d0);
        // kFast: SV / sqrt(16 log2 e) (or sqrt(512 log2 e)), so that its reciprocal carries the scale of the exp argument
        // (the weights a.xw are divided by the same constant on the host)
        r[8 + c] = C::kFast ? m.sv[i] * (1.0 / C::kScale) : m.sv[i];
        if (c == 0) {
            r[12] = m.origin[cell];
            r[13] = 0.0;
        }
    }
}

template <class C>
__device__ __forceinline__ void scale_cell_records(const ModelDev &m, const double *scal, double *rec, int tid, int nthr) {
    constexpr int kR = C::kRecD;
    const double amp = C::kDisp ? scal[5] : scal[4];
    for (int i = tid; i < m.ncell * 4; i += nthr) {
        double *r = rec + (i >> 2) * kR;
        const int c = i & 3;
        r[4 + c] = amp * r[4 + c];
        if (C::kDisp) r[C::kD1 + c] = (c == 0) ? fma(amp, r[C::kD1 + c], 1.0) : amp * r[C::kD1 + c];
    }
}

// dispersion model: first guess of the coordinate map, the velocity at the redshift-space point itself (:660)
template <class C>
__device__ __forceinline__ void first_guess(QuadCtx &q) {
    q.first = q.ifirst = 0.0;
    if (!C::kDisp) return;
    double S, iS;
    const double S2 = fma(q.Spar, q.Spar, q.Sperp2);
    if (C::kFast) {
        iS = fast_rsqrt(S2);
        S = S2 * iS;
    } else {
        S = sqrt(S2);
        iS = 1.0 / S;
    }
    const unsigned r0 = cell_record<C>(q, S);
    const double gv = cubic_at(r0 + 32, cell_coord(r0, S));
    q.first = C::kFast ? fma(gv, iS, 1.0) : 1.0 + gv / S;
    q.ifirst = C::kFast ? rcp_cubic(q.first) : 0.0;
}

// kMinBlocks = 4: 64 registers per thread -> 4 resident blocks of 256 threads per SM
// kFuse: the block also turns its row's theory vector into chi2 / lnL (k2_chi2.cuh: block_chi2); a
// separate instantiation, so the plain kernel's schedule is untouched by the epilogue
template <class C, bool kFuse = false>
__global__ void __launch_bounds__(256, C::kMinBlocks) k_multipoles(const __grid_constant__ K1Args a) {
    constexpr int kU = C::kU;
    constexpr int kR = C::kRecD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ModelDev &m = a.m;
    const int ncell = m.ncell, nx = m.nx;
    double *rec = reinterpret_cast<double *>(smem_raw);
    double *etab = rec + (size_t)ncell * kR;
    double *stage = etab + C::kTab;
    double *scal = stage + (size_t)a.jper * a.nmu;
    const int fitd = kFuse ? fused_fit_doubles(a.f.p) : 0;
    double *upper = scal + kNScal + fitd;
    int *bbase = reinterpret_cast<int *>(upper + ncell);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const long long row = blockIdx.x / a.nsplit;
    const int split = blockIdx.x - (int)(row * a.nsplit);
    const int j0 = split * a.jper;
    const int jn = min(a.jper, a.ns - j0);
    if (jn <= 0) return;

    const double *pr = a.params + row * kNPar;
    const double beta = m.beta_dependent ? pr[1] : m.beta_fixed;

    // ---- per-row scalars (ccf_model.py:589-613; velocity amplitude :419, :435, :449) ----
    row_scalars_to_shared(m, pr, scal, tid);
    for (int i = tid; i < ncell; i += nthr) upper[i] = m.upper[i];
    for (int i = tid; i < m.nbucket; i += nthr) bbase[i] = m.bucket_base[i];
    {
        const double *src = C::kBigTab ? m.exp_tab_big : m.exp_tab;
        for (int i = tid; i < C::kTab; i += nthr) etab[i] = src[i];
    }
    __syncthreads();
    build_cell_records<C>(m, scal, beta, rec, tid, nthr);
    __syncthreads();

    // ---- quadrature: thread <-> (s_j, mu_k), loop over the velocity nodes in registers ----
    const double sperp_f = scal[1], spar_f = scal[2];
    const int nmu = a.nmu;
    const int npairs = jn * nmu;
    QuadCtx q;
    q.kappa = scal[3];
    q.inv_h = pin_f64(m.inv_h);
    q.nbm1 = (unsigned)(m.nbucket - 1);
    q.rec_s = pin_u32((unsigned)__cvta_generic_to_shared(rec));
    q.etab_s = pin_u32((unsigned)__cvta_generic_to_shared(etab));
    q.bb_s = pin_u32((unsigned)__cvta_generic_to_shared(bbase));
    q.upper = upper;
    q.maxscan = m.maxscan;
    q.niter = m.niter;
    q.f_over_apar = scal[0] / scal[6];
    q.rt2 = 0.0;
    q.first = q.ifirst = 0.0;
    for (int pidx = tid; pidx < npairs; pidx += nthr) {
        const int jl = pidx / nmu, k = pidx - jl * nmu;
        const double sj = a.s[j0 + jl];
        const int km = a.pairwise ? j0 + jl : k;
        const double Sperp = sj * a.sqmu[km] * sperp_f;
        q.Spar = sj * a.mu[km] * spar_f;
        q.Sperp2 = Sperp * Sperp;
        if (C::kFromData) {
            const double rt = sj * a.sqmu[km];               // s_perp / aperp in fiducial units (:676)
            q.rt2 = rt * rt;
        }
        first_guess<C>(q);
        double acc = 0.0;
        int mi = 0;
        for (; mi + kU <= nx; mi += kU) acc = nodes<C, kU>(a, q, mi, acc);
        if (VB200_TAIL2 && kU >= 4 && mi + 2 <= nx) {   // the last nodes as one pair instead of two single trips
            acc = nodes<C, 2>(a, q, mi, acc);
            mi += 2;
        }
        for (; mi < nx; ++mi) acc = nodes<C, 1>(a, q, mi, acc);
        stage[pidx] = acc - 1.0;  // ccf_model.py:690
    }
    __syncthreads();
    // (everything the epilogue needs is re-derived from the kernel arguments here, so that nothing
    // extra stays live in registers across the quadrature loop)
    double *th2 = reinterpret_cast<double *>(smem_raw) + (size_t)a.m.ncell * kR + C::kTab + (size_t)a.jper * a.nmu + kNScal;
    write_outputs(a, stage, row, j0, jn, tid, nthr, kFuse ? th2 : nullptr);
    if (kFuse) {   // one block per row: finish with chi2 and lnL (ccf_fit.py:349-354, 441-483)
        __syncthreads();
        const long long r2 = blockIdx.x;   // nsplit == 1
        block_chi2(a.f, a.params[r2 * kNPar + 1], th2, th2 + ((a.f.p + 1) & ~1), r2, a.chi2, a.lnl, threadIdx.x,
                   blockDim.x);
    }
}

}  // namespace vb200
