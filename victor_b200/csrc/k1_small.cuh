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

constexpr int kSmallRows = 4;     // calls of up to this many rows use the kernel below
constexpr int kSmallPairs = 16;   // (s_j, mu_k) pairs per block
constexpr int kSmallLanes = 16;   // lanes per pair

struct SmallArgs {
    double *xi_scratch;      // [n][ns][nmu]
    unsigned *tickets;       // [n], zero before the launch; the last block of a row resets its counter
    double *theory;          // [n][p] or null
};

__host__ __device__ inline size_t small_smem_bytes(int ncell, int nbucket, int p, int rec, int tab) {
    size_t d = (size_t)ncell * (rec + 1) + tab + kNScal + fused_fit_doubles(p);
    return d * sizeof(double) + (size_t)nbucket * sizeof(int);
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
    int *bbase = reinterpret_cast<int *>(upper + ncell);

    const int tid = threadIdx.x, nthr = blockDim.x;
    const int nchunk = (nmu + kSmallPairs - 1) / kSmallPairs;
    const int per_row = ns * nchunk;
    const long long row = blockIdx.x / per_row;
    const int rem = blockIdx.x - (int)(row * per_row);
    const int j = rem / nchunk, kc = rem - j * nchunk;

    const double *pr = a.params + row * kNPar;
    const double beta = m.beta_dependent ? pr[1] : m.beta_fixed;

    // ---- prologue: as in k_multipoles (row scalars, then this row's cell records) ----
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

    // ---- one (s_j, mu_k) pair per half-warp, its velocity nodes over the 16 lanes ----
    const int pair = tid / kSmallLanes, l16 = tid % kSmallLanes;
    const int k = kc * kSmallPairs + pair;
    const int npl = (nx + kSmallLanes - 1) / kSmallLanes;
    double acc = 0.0;
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
    double *xi_row = sm.xi_scratch + (size_t)row * ns * nmu;
    if (l16 == 0 && k < nmu) xi_row[j * nmu + k] = acc - 1.0;   // ccf_model.py:690

    // ---- ticket: the last block of this row finishes it ----
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(sm.tickets + row, 1u);
        is_last = (t == (unsigned)(per_row - 1));
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (tid == 0) sm.tickets[row] = 0;   // ready for the next launch (stream-ordered after this one)

    // projection onto the multipoles: one warp per s_j, lane-strided FMAs + xor butterfly, exactly as
    // write_outputs does (ccf_model.py:824-825, utils.py:45-56), then the theory vector l-major (:856-858)
    {
        const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
        for (int jl = warp; jl < ns; jl += nwarp) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;
            for (int kk = lane; kk < nmu; kk += 32) {
                const double v = __ldcg(xi_row + jl * nmu + kk);
                s0 = fma(a.wmu[kk], v, s0);
                if (a.L > 1) s1 = fma(a.wmu[nmu + kk], v, s1);
                if (a.L > 2) s2 = fma(a.wmu[2 * nmu + kk], v, s2);
            }
            s0 = warp_sum(s0);
            if (a.L > 1) s1 = warp_sum(s1);
            if (a.L > 2) s2 = warp_sum(s2);
            if (lane == 0) {
                th[jl] = s0;
                if (a.L > 1) th[ns + jl] = s1;
                if (a.L > 2) th[2 * ns + jl] = s2;
            }
        }
    }
    __syncthreads();
    if (sm.theory)
        for (int i = tid; i < a.f.p; i += nthr) sm.theory[(size_t)row * a.f.p + i] = th[i];
    __syncthreads();
    block_chi2(a.f, pr[1], th, th + ((a.f.p + 1) & ~1), row, a.chi2, a.lnl, tid, nthr);
}

}  // namespace vb200
