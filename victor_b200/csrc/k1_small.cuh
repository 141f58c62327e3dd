// K1 + K2 for a handful of parameter rows (an MCMC step: n = 1): one launch, the velocity nodes spread over
// lanes and reduced with warp shuffles, chi-square finished by the last block to retire.
//
// The batch kernel (k1_streaming.cuh) gives each thread a whole (s_j, mu_k) pair and walks its 50 velocity nodes
// one after the other -- right when there are thousands of rows, but a single row then occupies 30 blocks
// whose threads each run a 13-trip dependent chain (21 us), and a second launch (K2) follows.  Here
//   * a block owns ONE s_j and kSmallPairs = 16 values of mu_k; the 16 lanes of a half-warp share a pair and take
//     ceil(nx / 16) consecutive velocity nodes each (nx = 50: 4 nodes = one trip of the same quad_nodes /
//     disp_nodes bodies the batch kernel runs), so a row spreads over ns * ceil(nmu / 16) = 210 blocks;
//   * the 16 partial Simpson sums of a pair are added by a shuffle butterfly (xor 8, 4, 2, 1);
//   * xi(s_j, mu_k) goes to a global scratch row; every block then takes a ticket (atomicAdd after a
//     __threadfence); the block that draws the last ticket of its row projects the row's xi onto the multipoles
//     (same lane order as write_outputs) and runs block_chi2 -- no second launch, no theory round trip.
// Summation order: per-lane consecutive nodes, then the butterfly -- NOT the batch kernel's m = 0 .. nx-1 chain,
// so results differ from the batch path in the last bits (measured <= 2e-15 relative on the multipoles);
// tests/test_gpu_parity.py::test_small_row_kernel holds both to the same goldens and to each other.
// Replaces, for n <= kSmallRows: CCFLikelihood.calculate -> CCFFit.log_likelihood (ccf_fit.py:356-483) and
// everything below it, as one launch.
#pragma once
#include "k1_streaming.cuh"

namespace vb200 {

constexpr int kSmallRows = 2;     // calls of up to this many rows use the kernel below (measured: faster than K1 + K2
                                  // for n <= 2, slower from n = 3 on, profiles/r02l_small_call_latency.txt)
constexpr int kSmallPairs = 16;   // (s_j, mu_k) pairs per block
constexpr int kSmallLanes = 16;   // lanes per pair

#ifdef VB200_SMALL_TIMING   // diagnostic build: globaltimer stamps of block 0 and of the last block (tools/probe_small_phases.py)
#define VB_STAMP(slot) do { if (threadIdx.x == 0 && (blockIdx.x == 0 || (slot) >= 5)) { unsigned long long t__; \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__)); sm.stamps[(slot)] = t__; } } while (0)
#else
#define VB_STAMP(slot) do { } while (0)
#endif

struct SmallArgs {
#ifdef VB200_SMALL_TIMING
    unsigned long long *stamps;
#endif
    double *xi_scratch;      // [n][ns * nmu] doubles per row; holds the blocks' projection shares [L][ns][nchunk]
    unsigned *tickets;       // [n], zero before the launch; the last block of a row resets its counter
    double *theory;          // [n][p] or null
    unsigned *done;          // [n] or null: host-mapped flags, set to 1 (after a system-scope fence) once a row's
                             // chi2 / lnL have been written -- the host may poll them instead of synchronising
    int inline_rows;         // 1: the parameter rows travel in `rows` below (kernel-parameter constant bank) instead of
    double rows[kSmallRows][kNPar];   // a device buffer: no copy node in front of the kernel, no global load in its prologue
};

// the last block stages both precision matrices of its row's bracket in shared memory (asynchronous copies that
// run under the projection) when they fit: p <= 64 and even (16-byte rows)
__host__ __device__ inline bool small_stages_matrices(int p) { return p <= 64 && p % 2 == 0; }

__host__ __device__ inline size_t small_smem_bytes(int ncell, int nbucket, int p, int rec, int tab) {
    size_t d = (size_t)ncell * (rec + 1) + tab + kNScal + fused_fit_doubles(p);
    d = (d + 1) & ~(size_t)1;
    if (small_stages_matrices(p)) d += 2 * (size_t)p * p;
    return d * sizeof(double) + (size_t)nbucket * sizeof(int);
}

// cov_bracket of k2_chi2.cuh, grid over the lanes (nbeta_cov <= 32); same conventions, same results
__device__ __forceinline__ void cov_bracket_warp(const FitDev &f, double beta, int lane, int &lo, int &hi, double &w) {
    lo = hi = 0;
    w = 0.0;
    if (f.cov_fixed) return;
    const int nb = f.nbeta_cov;
    const double gl = lane < nb ? f.beta_cov[lane] : 0.0;
    const double g0 = __shfl_sync(0xffffffffu, gl, 0), glast = __shfl_sync(0xffffffffu, gl, nb - 1);
    const unsigned below_m = __ballot_sync(0xffffffffu, lane < nb && gl < beta);
    const unsigned exact_m = __ballot_sync(0xffffffffu, lane < nb && gl == beta);
    if (beta < g0) return;
    if (beta > glast) {
        lo = hi = nb - 1;
        return;
    }
    const int below = __popc(below_m);
    if (exact_m) {
        lo = hi = 31 - __clz(exact_m);
    } else if (below == 0) {  // beta is NaN: every comparison false
        w = beta;
    } else {
        lo = below - 1;
        hi = nb - 1;  // sic: last index with grid >= beta (ccf_fit.py:225-227)
        const double glo = __shfl_sync(0xffffffffu, gl, lo);
        w = (beta - glo) / (glast - glo);
    }
}

// The last block of a row: projection of xi(s, mu) onto the multipoles, residual, the two quadratic forms,
// log-det term, likelihood -- block_chi2's arithmetic in block_chi2's order (bit-identical chi-square for the same
// theory vector), arranged for LATENCY: everything that does not depend on the theory vector is requested first --
// the bracket from one round of lane-parallel loads, both precision matrices by asynchronous global-to-shared
// copies (cp.async, no registers held), data-vector table and generalised eigenvalues into a few registers --
// and arrives while the projection runs.  `mats`: 2 p^2 doubles of shared memory, 16-byte aligned.
__device__ __forceinline__ void small_epilogue(const K1Args &a, const SmallArgs &sm, const double *part_row, int nchunk,
                                               long long row, double beta, double *th, double *mats, int tid, int nthr) {
    const FitDev &f = a.f;
    const int p = f.p, ns = a.ns, nmu = a.nmu;
    const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
    double *red = th + ((p + 1) & ~1);
    const bool fast = small_stages_matrices(p) && nwarp == kK2Warps && f.nbeta_cov <= 32 && f.nbeta_ccf <= 33;
    int lo = 0, hi = 0, kd = 0;
    double w = 0.0, td = 0.0;
    double dcoef[4] = {0.0, 0.0, 0.0, 0.0}, lam0 = 1.0, lam1 = 1.0;
    if (fast) {
        cov_bracket_warp(f, beta, lane, lo, hi, w);
        {
            const double *Mlo = f.icov + (size_t)lo * p * p, *Mhi = f.icov + (size_t)hi * p * p;
            const unsigned dst = (unsigned)__cvta_generic_to_shared(mats);
            const int chunks = p * p / 2;                          // 16-byte pieces per matrix
            for (int i = tid; i < chunks; i += nthr) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(Mlo + 2 * i) : "memory");
                if (hi != lo)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (chunks + i) * 16), "l"(Mhi + 2 * i)
                                 : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        if (f.data_beta_dependent) {
            kd = beta_interval_warp(f.beta_ccf, f.nbeta_ccf, beta, lane);
            td = beta - f.beta_ccf[kd];
        }
        if (tid < p) {
            const double *dt = f.data_tab + (size_t)kd * 4 * p;
#pragma unroll
            for (int q = 0; q < 4; ++q) dcoef[q] = dt[q * p + tid];
        }
        if (warp == 0 && f.use_logdet && hi != lo) {
            const double *lam = f.lam + (size_t)lo * p;
            lam0 = lane < p ? lam[lane] : 1.0;
            lam1 = lane + 32 < p ? lam[lane + 32] : 1.0;
        }
    }
    // projection: the blocks of the row left their shares w_l xi of every s_j (see k_small); add them in block order
    // and lay the theory vector out l-major (ccf_model.py:856-858).  One round of loads.
    for (int i = tid; i < a.L * ns; i += nthr) {
        const double *pp = part_row + (size_t)i * nchunk;
        double sum = 0.0;
#pragma unroll 8
        for (int c = 0; c < nchunk; ++c) sum += __ldcg(pp + c);
        th[i] = sum;
    }
    __syncthreads();
    VB_STAMP(6);
    if (sm.theory)
        for (int i = tid; i < p; i += nthr) sm.theory[(size_t)row * p + i] = th[i];
    if (!fast) {
        __syncthreads();
        block_chi2(f, beta, th, red, row, a.chi2, a.lnl, tid, nthr);
        if (sm.done) {   // (block_chi2's lane 0 of warp 0 wrote the results: same thread, fence, then the flag)
            if (tid == 0) {
                __threadfence_system();
                *reinterpret_cast<volatile unsigned *>(sm.done + row) = 1u;
            }
        }
        return;
    }
    // residual against the beta-PCHIP data vector (ccf_fit.py:193, 322-323, 350): DataAt's Horner form
    if (tid < p) th[tid] -= fma(fma(fma(dcoef[3], td, dcoef[2]), td, dcoef[1]), td, dcoef[0]);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    {   // the warps take one partial sum each (part = warp), as in block_chi2
        const double qlo = quad_partial(mats, th, p, lane, warp);
        const double qhi = (hi != lo) ? quad_partial(mats + (size_t)p * p, th, p, lane, warp) : 0.0;
        if (lane == 0) {
            red[warp] = qlo;
            red[kK2Warps + warp] = qhi;
        }
    }
    __syncthreads();
    VB_STAMP(7);
    if (warp == 0) {
        double qa = 0.0, qb = 0.0;
        for (int i = 0; i < kK2Warps; ++i) {   // same order as quad_form
            qa += red[i];
            qb += red[kK2Warps + i];
        }
        const double chi2 = blend_chi2(qa, qb, lo, hi, w);
        double norm = 0.0;
        if (f.use_logdet) {                    // norm_term of k2_chi2.cuh: lanes j and j + 32, then the butterfly
            double ld = 0.0;
            if (hi != lo) {
                if (lane < p) ld += log1p(w * (lam0 - 1.0));
                if (lane + 32 < p) ld += log1p(w * (lam1 - 1.0));
                ld = warp_sum(ld);
            }
            norm = -0.5 * (f.logdet[lo] + ld);
        }
        if (lane == 0) {
            store_likelihood(f, chi2, norm, row, a.chi2, a.lnl);
            if (sm.done) {
                __threadfence_system();
                *reinterpret_cast<volatile unsigned *>(sm.done + row) = 1u;
            }
        }
    }
}

template <class C>
__global__ void __launch_bounds__(kSmallPairs * kSmallLanes) k_small(const __grid_constant__ K1Args a,
                                                                     const __grid_constant__ SmallArgs sm) {
    constexpr int kR = C::kRecD;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int is_last;
    const ModelDev &m = a.m;
    const int ncell = m.ncell, nx = m.nx, nmu = a.nmu, ns = a.ns;
    double *rec = reinterpret_cast<double *>(smem_raw);
    double *etab = rec + (size_t)ncell * kR;
    double *scal = etab + C::kTab;
    double *th = scal + kNScal;
    double *upper = th + fused_fit_doubles(a.f.p);
    // (records, tables, scalars, th, upper) rounded up to a 16-byte boundary, then the matrices, then the bucket table
    const size_t head = (((size_t)ncell * (kR + 1) + C::kTab + kNScal + fused_fit_doubles(a.f.p)) + 1) & ~(size_t)1;
    double *mats = reinterpret_cast<double *>(smem_raw) + head;
    int *bbase = reinterpret_cast<int *>(mats + (small_stages_matrices(a.f.p) ? 2 * (size_t)a.f.p * a.f.p : 0));

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int nchunk = (nmu + kSmallPairs - 1) / kSmallPairs;
    const int per_row = ns * nchunk;
    const long long row = blockIdx.x / per_row;
    const int rem = blockIdx.x - (int)(row * per_row);
    const int j = rem / nchunk, kc = rem - j * nchunk;

    const double *pr = sm.inline_rows ? sm.rows[row] : a.params + row * kNPar;
    const double beta = m.beta_dependent ? pr[1] : m.beta_fixed;

    // ---- prologue: as in k_multipoles (row scalars, then this row's cell records) ----
    VB_STAMP(0);
    // warp 0 works out the row scalars (a chain of divisions and a square-root trapezoid, ~2 us) while the other
    // warps fetch the tables and build the records without the velocity amplitude, which is applied afterwards
    const bool overlap = !m.vd_beta_dep && !m.v0b && nthr > 32;
    row_scalars_to_shared(m, pr, scal, tid);
    if (tid >= 32 || !overlap) {
        const int t2 = overlap ? tid - 32 : tid, n2 = overlap ? nthr - 32 : nthr;
        for (int i = t2; i < ncell; i += n2) upper[i] = m.upper[i];
        for (int i = t2; i < m.nbucket; i += n2) bbase[i] = m.bucket_base[i];
        const double *src = C::kBigTab ? m.exp_tab_big : m.exp_tab;
        for (int i = t2; i < C::kTab; i += n2) etab[i] = src[i];
        if (overlap) build_cell_records<C, true>(m, scal, beta, rec, t2, n2);
    }
    __syncthreads();
    VB_STAMP(1);
    if (overlap)
        scale_cell_records<C>(m, scal, rec, tid, nthr);
    else
        build_cell_records<C>(m, scal, beta, rec, tid, nthr);
    __syncthreads();
    VB_STAMP(2);

    // ---- one (s_j, mu_k) pair per half-warp, its velocity nodes over the 16 lanes ----
    const int pair = tid / kSmallLanes, l16 = tid % kSmallLanes;
    const int k = kc * kSmallPairs + pair;
    const int npl = (nx + kSmallLanes - 1) / kSmallLanes;
    double acc = 0.0;
    double wk[kMaxPoles];   // projection weights of this pair's mu_k: fetched now, used after the quadrature
#pragma unroll
    for (int l = 0; l < kMaxPoles; ++l) wk[l] = (l < a.L && k < nmu) ? a.wmu[l * nmu + k] : 0.0;
    if (k < nmu) {
        QuadCtx q;
        q.kappa = scal[3];
        q.inv_h = m.inv_h;
        q.nbm1 = (unsigned)(m.nbucket - 1);
        q.rec_s = (unsigned)__cvta_generic_to_shared(rec);
        q.etab_s = (unsigned)__cvta_generic_to_shared(etab);
        q.bb_s = (unsigned)__cvta_generic_to_shared(bbase);
        q.upper = upper;
        q.maxscan = m.maxscan;
        q.niter = m.niter;
        const double sj = a.s[j];
        const double Sperp = sj * a.sqmu[k] * scal[1];
        q.Spar = sj * a.mu[k] * scal[2];
        q.Sperp2 = Sperp * Sperp;
        q.f_over_apar = scal[0] / scal[6];
        {
            const double rt = sj * a.sqmu[k];
            q.rt2 = rt * rt;
        }
        first_guess<C>(q);
        int mi = l16 * npl;
        const int mend = min(mi + npl, nx);
        for (; mi + 4 <= mend; mi += 4) acc = nodes<C, 4>(a, q, mi, acc);
        if (mi + 2 <= mend) {
            acc = nodes<C, 2>(a, q, mi, acc);
            mi += 2;
        }
        if (mi < mend) acc = nodes<C, 1>(a, q, mi, acc);
    }
#pragma unroll
    for (int o = kSmallLanes / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    // This block's share of the mu projection (ccf_model.py:690, :824-825, utils.py:45-56): w_l(mu_k) xi(s_j, mu_k) of
    // its 16 pairs, added in pair order by one thread per multipole; the row's last block adds the shares of the
    // nchunk blocks of every s_j in block order.  (The batch kernels add the same products lane-strided + butterfly:
    // a few ulp of difference, like the node sums above.)
    __shared__ double wxi[kMaxPoles][kSmallPairs];
    if (l16 == 0) {
        const double xi = acc - 1.0;
#pragma unroll
        for (int l = 0; l < kMaxPoles; ++l)
            if (l < a.L) wxi[l][pair] = k < nmu ? wk[l] * xi : 0.0;
    }
    __syncthreads();
    double *part_row = sm.xi_scratch + (size_t)row * ns * nmu;   // [L][ns][nchunk] shares (L nchunk <= nmu)
    if (tid < a.L) {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < kSmallPairs; ++q) sum += wxi[tid][q];
        part_row[((size_t)tid * ns + j) * nchunk + kc] = sum;
    }

    // ---- ticket: the last block of this row finishes it ----
    __syncthreads();
    VB_STAMP(3);
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(sm.tickets + row, 1u);
        is_last = (t == (unsigned)(per_row - 1));
    }
    __syncthreads();
    VB_STAMP(4);
    if (!is_last) return;
    VB_STAMP(5);
    __threadfence();
    if (tid == 0) sm.tickets[row] = 0;   // ready for the next launch (stream-ordered after this one)

    small_epilogue(a, sm, part_row, nchunk, row, pr[1], th, mats, tid, nthr);
    VB_STAMP(8);
}

}  // namespace vb200
