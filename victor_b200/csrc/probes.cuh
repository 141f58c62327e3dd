// Self-test and measurement kernels: hand-rolled math against the host, FP64 issue-rate probes.
#pragma once
#include "common.cuh"

namespace vb200 {

// ---------------------------------------------------------------------------------------
// math self-test kernel
// ---------------------------------------------------------------------------------------
constexpr int kSelftestOutputs = 10;

__global__ void k_math_selftest(const double *x, long long n, const double *etab, double *out) {
    __shared__ double tab[kExpTab];
    __shared__ double big[kExpTabBig];
    if (threadIdx.x < kExpTab) tab[threadIdx.x] = etab[threadIdx.x];
    for (int j = threadIdx.x; j < kExpTabBig; j += blockDim.x) big[j] = etab[kExpTab + j];
    __syncthreads();
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);
    out[i] = gauss_tab<6>(v, tab_s);   // exp(-v/2)
    out[n + i] = fast_rsqrt(v);
    out[2 * n + i] = rcp_cubic(v);
    out[3 * n + i] = gauss_tab<5>(v, tab_s);
    // the scaled-argument forms the tuned kernel uses: zs = sqrt(v) sqrt(16 log2 e)
    const double zs = sqrt(v) * kGaussScale;
    out[4 * n + i] = gauss_tab_scaled<52>(zs, tab_s);
    out[5 * n + i] = gauss_tab_scaled<53>(zs, tab_s);
    double u, mur;
    radius<2>(v, 1.0, u, mur);         // one Newton step: mur = 1 / sqrt(v)
    out[6 * n + i] = mur;
    out[7 * n + i] = rcp_newton(v);
    // 1024-entry table, degree-3 remainder ("exp_degree" 3: conversion-unit range reduction; 30: magic-number FMA)
    const unsigned big_s = (unsigned)__cvta_generic_to_shared(big);
    const double zb = sqrt(v) * kGaussScaleBig;
    out[8 * n + i] = gauss_big<true>(zb, big_s);
    out[9 * n + i] = gauss_big<false>(zb, big_s);
}

// ---------------------------------------------------------------------------------------
// pipe probe: which issue pipe do the FP64 <-> int/float conversions use, and how exact are the
// MUFU seeds?  mode 0: DFMA only; 1: cvt.rmi.s32.f64 only; 2: both interleaved 1:1;
// 3: cvt.rn.f32.f64 only; 4: DFMA + cvt.rn.f32.f64; 5: I2F (s32 -> f64) only; 6: DFMA + I2F.
// ---------------------------------------------------------------------------------------
template <int kMode>
__global__ void __launch_bounds__(256) k_pipe_probe(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3 + 1.0, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double y0 = x0 * 0.5, y1 = x1 * 0.5, y2 = x2 * 0.5, y3 = x3 * 0.5;
    int i0 = 0, i1 = 0, i2 = 0, i3 = 0;
    float f0 = 0.f, f1 = 0.f, f2 = 0.f, f3 = 0.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (kMode == 0 || kMode == 2 || kMode == 4 || kMode == 6) {
                x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            }
            if (kMode == 1 || kMode == 2) {
                int t0, t1, t2, t3;
                asm volatile("cvt.rmi.s32.f64 %0, %1;" : "=r"(t0) : "d"(y0));
                asm volatile("cvt.rmi.s32.f64 %0, %1;" : "=r"(t1) : "d"(y1));
                asm volatile("cvt.rmi.s32.f64 %0, %1;" : "=r"(t2) : "d"(y2));
                asm volatile("cvt.rmi.s32.f64 %0, %1;" : "=r"(t3) : "d"(y3));
                i0 ^= t0; i1 ^= t1; i2 ^= t2; i3 ^= t3;
                // feed the result back into the low word of the next input (integer ops only)
                y0 = __hiloint2double(__double2hiint(y0), i0 + r); y1 = __hiloint2double(__double2hiint(y1), i1 + r);
                y2 = __hiloint2double(__double2hiint(y2), i2 + r); y3 = __hiloint2double(__double2hiint(y3), i3 + r);
            }
            if (kMode == 3 || kMode == 4) {
                float t0, t1, t2, t3;
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t0) : "d"(y0));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t1) : "d"(y1));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t2) : "d"(y2));
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(t3) : "d"(y3));
                f0 += t0; f1 += t1; f2 += t2; f3 += t3;
                y0 = __hiloint2double(__double2hiint(y0), __float_as_int(f0)); y1 = __hiloint2double(__double2hiint(y1), __float_as_int(f1));
                y2 = __hiloint2double(__double2hiint(y2), __float_as_int(f2)); y3 = __hiloint2double(__double2hiint(y3), __float_as_int(f3));
            }
            if (kMode == 5 || kMode == 6) {
                double t0, t1, t2, t3;
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t0) : "r"(i0 + r));
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t1) : "r"(i1 + r));
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t2) : "r"(i2 + r));
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(t3) : "r"(i3 + r));
                i0 ^= __double2loint(t0); i1 ^= __double2loint(t1); i2 ^= __double2loint(t2); i3 ^= __double2loint(t3);
            }
        }
    }
    double s = (x0 + x1) + (x2 + x3) + (double)(i0 + i1 + i2 + i3) + (double)(f0 + f1 + f2 + f3);
    if (s == 12345.678) out[0] = s;
}

// DFMA issue model probe: kChains independent dependent-FMA chains per thread, with kMix other
// instructions after every DFMA (kKind 0: integer IMAD, 1: LDS.64 broadcast, 2: DMUL with a
// constant-bank operand instead of the plain DFMA).  128 threads per block = 1 warp per SMSP.
template <int kChains, int kMix, int kKind>
__global__ void __launch_bounds__(128) k_mix_probe(double *out, int iters, double a, double b, int ia) {
    __shared__ double sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 1.0 + threadIdx.x * 1e-9;
    __syncthreads();
    double x[kChains];
    int n[kChains];
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        x[c] = threadIdx.x * 1e-3 + c;
        n[c] = threadIdx.x + c;
    }
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sh);
    if (kKind == 5 || kKind == 6 || kKind == 7) {
        // register-operand bandwidth: kind 5 = DFMA with three distinct register operands per chain,
        // kind 6 = two distinct register operands + an immediate, kind 7 = DMUL of two registers
        double y[kChains], z[kChains];
#pragma unroll
        for (int c = 0; c < kChains; ++c) {   // values the compiler cannot relate to each other
            y[c] = out[1 + 2 * c + (threadIdx.x & 1)];
            z[c] = out[33 + 2 * c + (threadIdx.x & 1)];
        }
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int c = 0; c < kChains; ++c) {
                    if (kKind == 5) x[c] = fma(x[c], y[c], z[c]);
                    else if (kKind == 6) x[c] = fma(x[c], y[c], 1e-9);
                    else x[c] = x[c] * y[c];
                }
            }
        }
#pragma unroll
        for (int c = 0; c < kChains; ++c) acc += y[c] + z[c];
    } else if (kKind == 3 || kKind == 4) {
        // kind 3: the chains one after the other, 6 dependent DFMAs each (how ptxas orders the exp
        // polynomial of an unrolled loop); kind 4: the same work in interleaved order.  volatile asm
        // pins the order.
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                if (kKind == 3) {
#pragma unroll
                    for (int c = 0; c < kChains; ++c)
#pragma unroll
                        for (int k = 0; k < 6; ++k)
                            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b));
                } else {
#pragma unroll
                    for (int k = 0; k < 6; ++k)
#pragma unroll
                        for (int c = 0; c < kChains; ++c)
                            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[c]) : "d"(a), "d"(b));
                }
            }
        }
    } else
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) {
                if (kKind == 2) x[c] = x[c] * kExpPoly[0] + b; else x[c] = fma(x[c], a, b);
#pragma unroll
                for (int k = 0; k < kMix; ++k) {
                    if (kKind == 1) {
                        n[c] += __double2loint(lds_f64(sbase + ((n[c] & 7) << 3))) + ia;
                    } else {
                        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(n[c]) : "r"(ia), "r"(k));
                    }
                }
            }
        }
    }
    double s = acc;
    int t = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        s += x[c];
        t += n[c];
    }
    if (s == 12345.678 || t == 123456789) out[0] = s + t;
}

// Load-path probe: a cubic's four coefficients per lane and evaluation, fetched either from shared memory
// (2 x LDS.128, the kernels' way) or from the kernel-parameter constant bank with a per-lane index (4 x LDC.64),
// followed by the three Horner DFMAs.  `distinct` = different cells among the 32 lanes of a warp (1 = broadcast).
// Answers: is the constant path a usable second source of per-lane table data beside the shared-memory crossbar?
struct LoadProbeArgs {
    double tab[128 * 4];   // 128 cells x 4 coefficients = 4 KB of kernel parameters (constant bank 0)
    double *out;
    int iters, distinct, stride;
};

template <int kPath>
__global__ void __launch_bounds__(256) k_load_probe(const __grid_constant__ LoadProbeArgs a) {
    __shared__ __align__(16) double sh[128 * 4];
    for (int i = threadIdx.x; i < 128 * 4; i += blockDim.x) sh[i] = a.tab[i];
    __syncthreads();
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sh);
    const int lane = threadIdx.x & 31;
    int cell = (lane * a.distinct) >> 5;
    double t = 1e-3 * threadIdx.x, acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    for (int i = 0; i < a.iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            double c0, c1, c2, c3;
            const int c = (cell + r * a.stride) & 127;
            if (kPath == 0) {
                const double2 c01 = lds_f64x2(sbase + c * 32), c23 = lds_f64x2(sbase + c * 32 + 16);
                c0 = c01.x; c1 = c01.y; c2 = c23.x; c3 = c23.y;
            } else {
                c0 = a.tab[c * 4]; c1 = a.tab[c * 4 + 1]; c2 = a.tab[c * 4 + 2]; c3 = a.tab[c * 4 + 3];
            }
            const double v = fma(fma(fma(c3, t, c2), t, c1), t, c0);
            if (r == 0) acc0 += v; else if (r == 1) acc1 += v; else if (r == 2) acc2 += v; else acc3 += v;
        }
        cell = (cell + 1) & 127;
    }
    const double s = (acc0 + acc1) + (acc2 + acc3);
    if (s == 12345.678) a.out[0] = s;
}

// ---------------------------------------------------------------------------------------
// Quadratic-form probe: q[i] = r_i^T P r_i for n rows r_i [64] (p = 60 zero-padded to 64) against ONE matrix
// P [64][64] -- the half of K2's work that all rows share (the upper bracket is always the last precision matrix,
// ccf_fit.py:225-227).  Two arrangements, same inputs:
//   kind 0  warp-tiled FMA: the production arrangement of k_chi2 (matrix staged in shared memory, one warp per
//           row, lanes over columns, r_i broadcast from shared memory)
//   kind 1  FP64 DMMA: mma.sync.aligned.m8n8k4.f64, a warp owns an 8-row strip: Y = R_strip P as 8 column tiles x
//           16 k-steps, then q = rowsum(Y o R_strip) and a 4-lane butterfly
// north_star: "FP64 DMMA only if ncu shows it beats warp-tiled FMA".
// ---------------------------------------------------------------------------------------
constexpr int kQP = 64;

__global__ void __launch_bounds__(256) k_quad_fma(const double *R, const double *P, long long n, double *q) {
    __shared__ double sP[kQP * kQP];
    __shared__ double sr[8][kQP];
    for (int i = threadIdx.x; i < kQP * kQP; i += blockDim.x) sP[i] = P[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (long long row = (long long)blockIdx.x * 8 + warp; row < n; row += (long long)gridDim.x * 8) {
        __syncwarp();
        sr[warp][lane] = R[row * kQP + lane];
        sr[warp][lane + 32] = R[row * kQP + lane + 32];
        __syncwarp();
        double y0 = 0.0, y1 = 0.0;
#pragma unroll 8
        for (int i = 0; i < kQP; ++i) {
            const double ri = sr[warp][i];
            y0 = fma(sP[i * kQP + lane], ri, y0);
            y1 = fma(sP[i * kQP + lane + 32], ri, y1);
        }
        double acc = fma(y0, sr[warp][lane], y1 * sr[warp][lane + 32]);
        acc = warp_sum(acc);
        if (lane == 0) q[row] = acc;
    }
}

__global__ void __launch_bounds__(256) k_quad_dmma(const double *R, const double *P, long long n, double *q) {
    __shared__ double sP[kQP * (kQP + 1)];   // [k][n], padded rows: B fragments read a column of 4 k's per group
    for (int i = threadIdx.x; i < kQP * kQP; i += blockDim.x) sP[(i / kQP) * (kQP + 1) + (i % kQP)] = P[i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (long long strip = (long long)blockIdx.x * 8 + warp; strip * 8 < n; strip += (long long)gridDim.x * 8) {
        const long long row = strip * 8 + g;               // this thread's row of the strip (A and C fragments)
        const double *r = R + (row < n ? row : n - 1) * kQP;
        double c[8][2];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) c[nt][0] = c[nt][1] = 0.0;
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
            const double a0 = r[ks * 4 + t];                // A[g][t] of this k-step
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const double b0 = sP[(ks * 4 + t) * (kQP + 1) + nt * 8 + g];   // B[t][g]
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c[nt][0]), "+d"(c[nt][1]) : "d"(a0), "d"(b0));
            }
        }
        double acc = 0.0;                                   // C[g][2t], C[g][2t + 1] of every column tile
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc = fma(c[nt][0], r[nt * 8 + 2 * t], fma(c[nt][1], r[nt * 8 + 2 * t + 1], acc));
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        if (t == 0 && row < n) q[row] = acc;
    }
}

// raw MUFU seeds (no refinement): out[0..n) = rsqrt.approx(x), out[n..2n) = rcp.approx(x)
__global__ void k_seed_probe(const double *x, long long n, double *out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double y, q;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x[i]));
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(x[i]));
    out[i] = y;
    out[n + i] = q;
}

// ---------------------------------------------------------------------------------------
// FP64 issue-rate probe: 8 independent DFMA chains per thread (roofline denominator measured
// on the box the bench runs on; MEASURED_PEAKS.json has no FP64 figure)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) out[0] = s;  // keep the chains alive
}

}  // namespace vb200
