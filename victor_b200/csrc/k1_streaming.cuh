// K1 (tuned): xi(s, mu) and its multipoles for the Gaussian streaming model with an isotropic
// real-space correlation and an isotropic sigma_v(r) template -- the BOSS DR12 CMASS setup.
//
// One thread block per (parameter row, s-bin range).  Replaces
//   CCFModel.theory_xi streaming branch   victor/ccf_model.py:589-658, 681-690
//   CCFModel.theory_multipoles            victor/ccf_model.py:816-825 + victor/utils.py:45-56
//   CCFModel.theory_multipole_vector      victor/ccf_model.py:856-858
// The block first turns the host tables into its row's own cell table in shared memory
// (beta-Horner of the xi^r power table, velocity amplitude folded into V0), then every thread
// owns (s_j, mu_k) pairs and runs the velocity quadrature over x_m in registers; xi(s_j, mu_k)
// is staged in shared memory and projected onto the multipoles with warp shuffles.
#pragma once
#include "common.cuh"
#include "k2_chi2.cuh"

namespace vb200 {

// Shared memory (dynamic), see k1_smem_bytes():
//   rec[ncell][14]  per-row cell records: xi+1 (4) | B*V0 (4) | SV (4) | origin | pad   (16 B aligned;
//                   stride 112 B = 28 banks, so 8 consecutive cells tile the 32 banks exactly)
//   etab[32]        2^(j/32)
//   stage[jper*nmu] xi(s_j, mu_k) of this block
//   scal[kNScal]    per-row scalars
//   th[fitd]        fused likelihood epilogue only: theory / residual vector and per-warp partial sums
//   upper[ncell]    upper knot of each cell (slow path of the cell search only)
//   int bbase[nbucket]  bucket -> first cell; bit 31 set when a knot lies strictly inside the bucket
constexpr int kRec = 14;

__host__ __device__ inline size_t k1_smem_bytes(int ncell, int jper, int nmu, int nbucket, int fitd = 0) {
    size_t d = (size_t)ncell * (kRec + 1) + kExpTab + (size_t)jper * nmu + kNScal + fitd;
    return d * sizeof(double) + (size_t)nbucket * sizeof(int);
}

// Compile-time variant.
//   kFast  : hand-rolled rsqrt / rcp / exp (else CUDA libm).
//   kNewton: 3 = cubic-convergence refinement of the MUFU seeds (default), 2 = one Newton step.
//   kFlags : some bucket holds a knot in its interior, so the cell search may need the
//            comparison path (non-lattice knot sets).
//   kU     : velocity nodes processed together per loop trip (instruction-level parallelism; a warp
//            must keep >= 4 independent DFMAs in flight to reach the FP64 issue rate).
//   kExp   : degree of the exp remainder polynomial (6 = Taylor, 5 = economised).
template <bool kFast_, bool kFlags_, int kU_, int kExp_, int kNewton_ = 3>
struct K1Cfg {
    static constexpr bool kFast = kFast_, kFlags = kFlags_;
    static constexpr int kU = kU_, kExp = kExp_;
    static constexpr int kMath = kFast_ ? (kNewton_ == 2 ? 2 : 1) : 0;
};

// per-thread loop invariants of the quadrature
struct QuadCtx {
    double kappa, Spar, Sperp2, inv_h;
    unsigned nbm1, bb_s, rec_s, etab_s;
    const double *upper;
    int maxscan;
};

// U consecutive velocity nodes of one (s_j, mu_k) pair, written stage by stage so that the U
// dependency chains are interleaved in program order (DFMA latency 8.5 cycles, issue every 2.2).
template <class C, int U>
__device__ __forceinline__ double quad_nodes(const K1Args &a, const QuadCtx &q, int mi, double acc) {
    double xm[U], u[U], mur[U], t[U], rq[U], z2[U], g[U];   // kFast: z2 holds zs = z sqrt(16 log2 e), not z^2
    unsigned ra[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
        xm[i] = a.xw[mi + i];                                 // uniform: constant-bank read
        const double rp = fma(-xm[i], q.kappa, q.Spar);       // ccf_model.py:648-650
        const double u2 = fma(rp, rp, q.Sperp2);              // :651
        radius<C::kMath>(u2, rp, u[i], mur[i]);               // :651-652
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
        // bucket = floor(u * inv_h) through a round-down FMA onto 1.5 * 2^52 (u >= 0; NaN -> 0)
        const unsigned b = min((unsigned)__double2loint(__fma_rd(u[i], q.inv_h, 6755399441055744.0)), q.nbm1);
        int cell = lds_s32(q.bb_s + (b << 2));
        if (C::kFlags) {
            if (cell < 0) {  // a knot lies inside this bucket: finish the search by comparison
                cell &= ~kBucketFlag;
                for (int sc = 0; sc < q.maxscan; ++sc) cell += (u[i] >= q.upper[cell]) ? 1 : 0;
            }
        }
        ra[i] = q.rec_s + cell * (kRec * 8);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
        // t = max(t, 0): below the first knot every spline is its boundary value (ext=3), which is
        // the first cell's cubic at t = 0.  Done on the high word with an integer max: a negative
        // t becomes a positive denormal-sized number, i.e. 0 for the cubic.
        double tt = u[i] - lds_f64(ra[i] + 96);
        const int hi = __double2hiint(tt);
        t[i] = __hiloint2double(max(hi, 0), __double2loint(tt));
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
        const double2 c89 = lds_f64x2(ra[i] + 64), cab = lds_f64x2(ra[i] + 80);
        const double sv = fma(fma(fma(cab.y, t[i], cab.x), t[i], c89.y), t[i], c89.x);   // :654-655
        rq[i] = C::kMath == 1 ? rcp_cubic(sv) : (C::kMath == 2 ? rcp_newton(sv) : 1.0 / sv);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
        const double2 c45 = lds_f64x2(ra[i] + 32), c67 = lds_f64x2(ra[i] + 48);
        const double vb = fma(fma(fma(c67.y, t[i], c67.x), t[i], c45.y), t[i], c45.x);   // :635, :656
        const double z = fma(-vb, mur[i], xm[i]) * rq[i];   // kFast: rq = sqrt(16 log2 e) / SV (scaled table)
        z2[i] = C::kFast ? z : z * z;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) g[i] = C::kFast ? gauss_tab_scaled<C::kExp>(z2[i], q.etab_s) : exp(-0.5 * z2[i]);
#pragma unroll
    for (int i = 0; i < U; ++i) {
        const double2 c01 = lds_f64x2(ra[i]), c23 = lds_f64x2(ra[i] + 16);
        const double xi1 = fma(fma(fma(c23.y, t[i], c23.x), t[i], c01.y), t[i], c01.x);  // :621, :683
        acc = fma(a.xw[kMaxNx + mi + i] * (xi1 * rq[i]), g[i], acc);                      // :690
    }
    return acc;
}

// 64 registers per thread -> 4 resident blocks of 256 threads per SM
// kFuse: the block also turns its row's theory vector into chi2 / lnL (k2_chi2.cuh: block_chi2); a
// separate instantiation, so the plain kernel's schedule is untouched by the epilogue
template <class C, bool kFuse = false>
__global__ void __launch_bounds__(256, 4) k_multipoles(const __grid_constant__ K1Args a) {
    constexpr int kU = C::kU;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ModelDev &m = a.m;
    const int ncell = m.ncell, nx = m.nx;
    double *rec = reinterpret_cast<double *>(smem_raw);
    double *etab = rec + (size_t)ncell * kRec;
    double *stage = etab + kExpTab;
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
    if (tid < kExpTab) etab[tid] = m.exp_tab[tid];
    __syncthreads();
    // ---- this row's cell records: xi^r(u; beta) from the beta power table (+1 folded into c0),
    //      B * V0(u), SV(u), origin ----
    {
        const double B = scal[4];
        int kb = 0;
        double tb = 0.0;
        if (m.beta_dependent) {
            kb = beta_interval(m.beta_grid, m.nbeta, beta);
            tb = beta - m.beta_grid[kb];
        }
        const double *tab = m.xi_tab + (size_t)kb * 4 * ncell * 4;  // [q][cell][4], ell index 0
        const int per = ncell * 4;
        for (int i = tid; i < per; i += nthr) {
            const int cell = i >> 2, c = i & 3;
            double v = fma(fma(fma(tab[3 * per + i], tb, tab[2 * per + i]), tb, tab[per + i]), tb, tab[i]);
            if (c == 0) v += 1.0;
            double *r = rec + cell * kRec;
            r[c] = v;
            double v0 = m.v0[i];
            if (m.vd_beta_dep) {   // linear_bias matter model: V0 follows the monopole's beta dependence (ccf_model.py:358-370)
                const double *tv = m.v0 + (size_t)kb * 4 * per;
                v0 = fma(fma(fma(tv[3 * per + i], tb, tv[2 * per + i]), tb, tv[per + i]), tb, tv[i]);
            } else if (m.v0b) {    // empirical correction (1 + Av delta(r)) of the mean velocity (:451-455)
                v0 = fma(scal[8], m.v0b[i], v0);
            }
            r[4 + c] = B * v0;
            // kFast: SV / sqrt(16 log2 e), so that its reciprocal carries the scale of the exp argument
            // (the weights a.xw are divided by the same constant on the host)
            r[8 + c] = C::kFast ? m.sv[i] * (1.0 / kGaussScale) : m.sv[i];
            if (c == 0) {
                r[12] = m.origin[cell];
                r[13] = 0.0;
            }
        }
    }
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
    for (int pidx = tid; pidx < npairs; pidx += nthr) {
        const int jl = pidx / nmu, k = pidx - jl * nmu;
        const double sj = a.s[j0 + jl];
        const int km = a.pairwise ? j0 + jl : k;
        const double Sperp = sj * a.sqmu[km] * sperp_f;
        q.Spar = sj * a.mu[km] * spar_f;
        q.Sperp2 = Sperp * Sperp;
        double acc = 0.0;
        int mi = 0;
        for (; mi + kU <= nx; mi += kU) acc = quad_nodes<C, kU>(a, q, mi, acc);
        for (; mi < nx; ++mi) acc = quad_nodes<C, 1>(a, q, mi, acc);
        stage[pidx] = acc - 1.0;  // ccf_model.py:690
    }
    __syncthreads();
    // (everything the epilogue needs is re-derived from the kernel arguments here, so that nothing
    // extra stays live in registers across the quadrature loop)
    double *th2 = reinterpret_cast<double *>(smem_raw) + (size_t)a.m.ncell * kRec + kExpTab + (size_t)a.jper * a.nmu + kNScal;
    write_outputs(a, stage, row, j0, jn, tid, nthr, kFuse ? th2 : nullptr);
    if (kFuse) {   // one block per row: finish with chi2 and lnL (ccf_fit.py:349-354, 441-483)
        __syncthreads();
        const long long r2 = blockIdx.x;   // nsplit == 1
        block_chi2(a.f, a.params[r2 * kNPar + 1], th2, th2 + ((a.f.p + 1) & ~1), r2, a.chi2, a.lnl, threadIdx.x,
                   blockDim.x);
    }
}

}  // namespace vb200
