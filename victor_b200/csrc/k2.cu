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

cudaError_t launch_k2_kernels(const K2Args &a, long long n, int sm_count, cudaStream_t st) {
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
