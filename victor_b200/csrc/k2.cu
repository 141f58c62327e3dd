// k2.cu: kernel instantiations of one family (see kernels.h); compiled as its own translation unit.
#include "k2_chi2.cuh"
#include "kernels.h"

namespace vb200 {

__global__ void __launch_bounds__(kK2Warps * 32) k_chi2(const __grid_constant__ K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FitDev &f = a.f;
    const int p = f.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *res = reinterpret_cast<double *>(smem_raw) + (size_t)warp * p;
    double *staged = reinterpret_cast<double *>(smem_raw) + (size_t)kK2Warps * p;
    const int staged_idx = k2_stages(p) ? (f.cov_fixed ? 0 : f.nbeta_cov - 1) : -1;
    if (staged_idx >= 0) {
        const double *src = f.icov + (size_t)staged_idx * p * p;
        for (int i = threadIdx.x; i < p * p; i += blockDim.x) staged[i] = src[i];
    }
    __syncthreads();
    for (int rr = 0; rr < kK2RowsPerWarp; ++rr) {
        const long long row = ((long long)blockIdx.x * kK2Warps + warp) * kK2RowsPerWarp + rr;
        if (row >= a.n) break;
        const double beta = a.params[row * kNPar + 1];
        const double *th = a.theory + (size_t)row * p;

        const DataAt data(f, beta);
        __syncwarp();
        for (int j = lane; j < p; j += 32) res[j] = th[j] - data(j);
        __syncwarp();

        int lo, hi;
        double w;
        cov_bracket(f, beta, lo, hi, w);
        const double *Mlo = (lo == staged_idx) ? staged : f.icov + (size_t)lo * p * p;
        const double *Mhi = (hi == staged_idx) ? staged : f.icov + (size_t)hi * p * p;
        const double qlo = quad_form(Mlo, res, p, lane);
        const double qhi = (hi != lo) ? quad_form(Mhi, res, p, lane) : 0.0;
        const double chi2 = blend_chi2(qlo, qhi, lo, hi, w);
        const double norm = norm_term(f, lo, hi, w, lane);
        if (lane == 0) store_likelihood(f, chi2, norm, row, a.chi2, a.lnl);
    }
}

// K2 for small batches (MCMC steps): one block per parameter row, the warps of the block sharing the
// rows of the precision matrices (block_chi2) -- a single warp walking 2 x p matrix rows out of L2 takes
// ~40 us per row, eight warps ~5 us.  Same partition of the sums as k_chi2: bit-identical results.
__global__ void __launch_bounds__(kK2Warps * 32) k_chi2_block(const __grid_constant__ K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const FitDev &f = a.f;
    const int p = f.p;
    double *th = reinterpret_cast<double *>(smem_raw);
    const long long row = blockIdx.x;
    for (int j = threadIdx.x; j < p; j += blockDim.x) th[j] = a.theory[(size_t)row * p + j];
    __syncthreads();
    block_chi2(f, a.params[row * kNPar + 1], th, th + ((p + 1) & ~1), row, a.chi2, a.lnl, threadIdx.x, blockDim.x);
}

// ---- bucketed K2 (large batches): rows grouped by lower covariance bracket -----------------------------------
// k_chi2 streams every row's own lower-bracket matrix out of L2 (1.5 GB per 65,536 rows).  Here the rows are first
// grouped by that bracket (two small kernels: count, scatter), and a block then serves kK2TileRows rows of one
// bracket from shared memory, two rows at a time against each fetched matrix element.  Per-row arithmetic and its
// order are k_chi2's: the results are bit-identical, whatever order the scatter happens to produce.
__global__ void __launch_bounds__(256) k_bracket_count(const __grid_constant__ K2Args a) {
    __shared__ unsigned h[kK2MaxBins];
    if (threadIdx.x < kK2MaxBins) h[threadIdx.x] = 0u;
    __syncthreads();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row < a.n) {
        int lo, hi;
        double w;
        cov_bracket(a.f, a.params[row * kNPar + 1], lo, hi, w);
        a.lo8[row] = (unsigned char)lo;
        atomicAdd(&h[lo], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kK2MaxBins && h[threadIdx.x]) atomicAdd(a.bins + threadIdx.x, h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_bracket_scatter(const __grid_constant__ K2Args a) {
    __shared__ unsigned h[kK2MaxBins], base[kK2MaxBins];
    if (threadIdx.x < kK2MaxBins) h[threadIdx.x] = 0u;
    __syncthreads();
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int lo = 0;
    unsigned mine = 0u;
    if (row < a.n) {
        lo = a.lo8[row];
        mine = atomicAdd(&h[lo], 1u);
    }
    __syncthreads();
    if (threadIdx.x < kK2MaxBins) {
        const int b = threadIdx.x;
        unsigned start = 0u;
        for (int i = 0; i < b; ++i) start += a.bins[i];                       // rows of the brackets in front
        base[b] = start + (h[b] ? atomicAdd(a.bins + kK2MaxBins + b, h[b]) : 0u);   // + this block's run inside bracket b
    }
    __syncthreads();
    if (row < a.n) a.order[base[lo] + mine] = (int)row;
}

// the four rows of one warp: residuals, both quadratic forms, normalisation, store
template <int kC>
__device__ __forceinline__ void bucket_rows(const K2Args &a, const double *Mlo, const double *Mhi, const double *lam_s,
                                            double *res4, const long long (&rows)[4], const double (&betas)[4], int valid,
                                            int b, int lane) {
    const FitDev &f = a.f;
    const int p = f.p;
    double own[4][kC];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const double *th = a.theory + (size_t)rows[r] * p;
        const DataAt data(f, betas[r], lane);
#pragma unroll
        for (int c = 0; c < kC; ++c) {
            const int j = lane + 32 * c;
            own[r][c] = j < p ? th[j] - data(j) : 0.0;
        }
    }
#pragma unroll
    for (int c = 0; c < kC; ++c) {
        const int j = lane + 32 * c;
        if (j < p) {
            *reinterpret_cast<double2 *>(res4 + 4 * j) = make_double2(own[0][c], own[1][c]);
            *reinterpret_cast<double2 *>(res4 + 4 * j + 2) = make_double2(own[2][c], own[3][c]);
        }
    }
    __syncwarp();
    // the rows' lower bracket is the tile's (that is how they were grouped): hi and w follow from three grid values
    const int last = f.cov_fixed ? 0 : f.nbeta_cov - 1;
    const double g0 = f.cov_fixed ? 0.0 : f.beta_cov[0], gb = f.cov_fixed ? 0.0 : f.beta_cov[b],
                 glast = f.cov_fixed ? 0.0 : f.beta_cov[last];
    int hi[4];
    double w[4];
    bool blend = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        bracket_from_lo(f, betas[r], b, g0, gb, glast, hi[r], w[r]);
        blend = blend || hi[r] != b;
    }
    const unsigned res_s = (unsigned)__cvta_generic_to_shared(res4);
    double qlo[4], qhi[4] = {0.0, 0.0, 0.0, 0.0};
    quad_form4_t<kC>((unsigned)__cvta_generic_to_shared(Mlo), res_s, own, p, lane, qlo);
    if (blend) quad_form4_t<kC>((unsigned)__cvta_generic_to_shared(Mhi), res_s, own, p, lane, qhi);
    const double ldb = f.use_logdet ? f.logdet[b] : 0.0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double norm = 0.0;
        if (f.use_logdet) {   // norm_term with the bracket's eigenvalue row out of shared memory (same sums, same order)
            double ld = 0.0;
            if (hi[r] != b) {
                for (int j = lane; j < p; j += 32) ld += log1p(w[r] * (lam_s[j] - 1.0));
                ld = warp_sum(ld);
            }
            norm = -0.5 * (ldb + ld);
        }
        if (lane == 0 && r < valid)
            store_likelihood(f, blend_chi2(qlo[r], qhi[r], b, hi[r], w[r]), norm, rows[r], a.chi2, a.lnl);
    }
}

__global__ void __launch_bounds__(kK2Warps * 32, 3) k_chi2_bucketed(const __grid_constant__ K2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int tile_info[3];   // bracket, first position in `order`, rows in this tile
    __shared__ unsigned bins_s[kK2MaxBins];
    static_assert(kK2TileRows == 4 * kK2Warps, "a warp takes four rows");
    const FitDev &f = a.f;
    const int p = f.p, pe = (p + 1) & ~1;
    const size_t pp = ((size_t)p * p + 1) & ~(size_t)1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *Mlo = reinterpret_cast<double *>(smem_raw);
    double *Mhi = Mlo + pp;
    double *lam_s = Mhi + pp;
    double *res4 = lam_s + pe + (size_t)warp * 4 * pe;
    const int nb = f.cov_fixed ? 1 : f.nbeta_cov;
    if (threadIdx.x < kK2MaxBins) bins_s[threadIdx.x] = threadIdx.x < nb ? a.bins[threadIdx.x] : 0u;   // one round of loads
    __syncthreads();
    if (threadIdx.x == 0) {   // tile -> (bracket, offset): brackets own ceil(count / kK2TileRows) consecutive tiles
        int tile = blockIdx.x, b = 0, first = 0, cnt = 0;
        for (; b < nb; ++b) {
            cnt = (int)bins_s[b];
            const int tiles = (cnt + kK2TileRows - 1) / kK2TileRows;
            if (tile < tiles) break;
            tile -= tiles;
            first += cnt;
        }
        tile_info[0] = b < nb ? b : -1;
        tile_info[1] = first + tile * kK2TileRows;
        tile_info[2] = b < nb ? min(kK2TileRows, cnt - tile * kK2TileRows) : 0;
    }
    __syncthreads();
    const int b = tile_info[0], first = tile_info[1], nrows = tile_info[2];
    if (b < 0) return;   // (the grid is sized for the worst case: every bracket with a partial tile)
    const int last = nb - 1;
    {   // both matrices (and the bracket's eigenvalue row) into shared memory, all copies in flight at once
        const double *slo = f.icov + (size_t)b * p * p, *shi = f.icov + (size_t)last * p * p;
        if ((p & 1) == 0) {
            const unsigned dlo = (unsigned)__cvta_generic_to_shared(Mlo), dhi = (unsigned)__cvta_generic_to_shared(Mhi);
            for (int i = threadIdx.x; i < p * p / 2; i += blockDim.x) {
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dlo + i * 16), "l"(slo + 2 * i) : "memory");
                if (b != last)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dhi + i * 16), "l"(shi + 2 * i) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            for (int i = threadIdx.x; i < p * p; i += blockDim.x) {
                Mlo[i] = slo[i];
                if (b != last) Mhi[i] = shi[i];
            }
        }
        if (f.use_logdet && b != last)   // (rows of the last bracket never blend: norm_term does not touch lam for them)
            for (int j = threadIdx.x; j < p; j += blockDim.x) lam_s[j] = f.lam[(size_t)b * p + j];
    }
    long long rows[4];
    double betas[4];
    const int valid = min(4, nrows - warp * 4);   // (<= 0: this warp has no rows in a short last tile)
#pragma unroll
    for (int k = 0; k < 4; ++k) rows[k] = a.order[first + min(warp * 4 + k, nrows - 1)];   // short: the last row again
#pragma unroll
    for (int k = 0; k < 4; ++k) betas[k] = a.params[rows[k] * kNPar + 1];
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (valid <= 0) return;
    switch ((p + 31) >> 5) {
        case 1: bucket_rows<1>(a, Mlo, Mhi, lam_s, res4, rows, betas, valid, b, lane); break;
        default: bucket_rows<2>(a, Mlo, Mhi, lam_s, res4, rows, betas, valid, b, lane); break;   // p <= kK2StageMaxP = 64
    }
}

cudaError_t launch_k2_kernels(const K2Args &a, long long n, int sm_count, cudaStream_t st) {
    if (a.order) {
        const int nb = a.f.cov_fixed ? 1 : a.f.nbeta_cov;
        const size_t smem = k2_bucket_smem_bytes(a.f.p);
        cudaError_t e = cudaFuncSetAttribute(k_chi2_bucketed, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess)   // three blocks of 73 KB per SM need the whole shared-memory carve-out
            e = cudaFuncSetAttribute(k_chi2_bucketed, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        e = cudaMemsetAsync(a.bins, 0, 2 * kK2MaxBins * sizeof(unsigned), st);
        if (e != cudaSuccess) return e;
        const unsigned gb = (unsigned)((n + 255) / 256);
        k_bracket_count<<<gb, 256, 0, st>>>(a);
        k_bracket_scatter<<<gb, 256, 0, st>>>(a);
        const long long tiles = (n + kK2TileRows - 1) / kK2TileRows + nb;
        k_chi2_bucketed<<<(unsigned)tiles, kK2Warps * 32, smem, st>>>(a);
        return cudaGetLastError();
    }
    if (n <= (long long)sm_count * 4) {
        // few rows: a block per row, so that one row's matrix reads are spread over eight warps
        k_chi2_block<<<(unsigned)n, kK2Warps * 32, (size_t)fused_fit_doubles(a.f.p) * sizeof(double), st>>>(a);
    } else {
        const long long rows_per_block = (long long)kK2Warps * kK2RowsPerWarp;
        const long long blocks = (n + rows_per_block - 1) / rows_per_block;
        k_chi2<<<(unsigned)blocks, kK2Warps * 32, k2_smem_bytes(a.f.p), st>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace vb200
