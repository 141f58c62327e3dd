/*
 * victor_b200_probes.h - measurement and self-test entry points of libvictor_b200_probes.so.
 *
 * NOT part of the product ABI (include/victor_b200.h): nothing a binding of the reference would call.
 * The library is built from the same device math as the kernels (victor_b200/csrc/common.cuh) and is
 * used by tests/ (accuracy of the hand-rolled exp / rsqrt / reciprocal), bench.py (FP64 issue-rate
 * figure reported beside the nominal peak) and tools/probe_*.py (the issue model of DESIGN.md section 5).
 * Every function returns 0 or a negative code; the message is in vb200p_last_error() (thread-local).
 */
#ifndef VICTOR_B200_PROBES_H
#define VICTOR_B200_PROBES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *vb200p_last_error(void);

/* Device self-test of the hand-rolled math for n HOST inputs x > 0; out has 10n entries:
 *   [0,n)   exp(-x/2), degree-6 remainder polynomial      [n,2n)  1/sqrt(x), cubic refinement
 *   [2n,3n) 1/x, cubic refinement                         [3n,4n) exp(-x/2), degree-5 economised
 *   [4n,5n) exp(-x/2), scaled argument, FP32 tail of two Horner steps ("exp_degree" 52)
 *   [5n,6n) the same with three FP32 steps ("exp_degree" 53)
 *   [6n,7n) 1/sqrt(x), one Newton step ("newton" 2)       [7n,8n) 1/x, one Newton step
 *   [8n,9n) exp(-x/2), 1024-entry table + degree-3 remainder, conversion-unit range reduction ("exp_degree" 3)
 *   [9n,10n) the same with the magic-number range reduction ("exp_degree" 30) */
int vb200p_math_selftest(int device, const double *x, int64_t n, double *out);

/* Time of a kernel issuing only DFMA (mode 0), only one kind of FP64 conversion (1, 3, 5) or both
 * interleaved (2, 4, 6); and the raw MUFU seeds rsqrt.approx / rcp.approx for n HOST inputs
 * (out has 2n entries). */
int vb200p_pipe_probe(int device, int mode, int iters, double *ms);
int vb200p_seed_probe(int device, const double *x, int64_t n, double *out);

/* DFMA issue-model probe: `chains` dependent FMA chains per thread, `mix` other instructions
 * (kind 0 integer, 1 shared-memory load; kind 2 = DMUL-with-constant-operand chains) after every
 * DFMA, `blocks_per_sm` blocks of 128 threads (= warps per SM sub-partition). */
int vb200p_mix_probe(int device, int chains, int mix, int kind, int blocks_per_sm, int iters, double *ms);

/* Load-path probe: 4 coefficient fetches per lane and cubic from shared memory (path 0: 2 x LDS.128) or from the
 * kernel-parameter constant bank with a per-lane index (path 1: 4 x LDC.64), + 3 DFMA; `distinct` = different cells
 * among the lanes of a warp, `stride` = cell distance between the four cubics of one trip; blocks of 256 threads. */
int vb200p_load_probe(int device, int path, int distinct, int stride, int blocks_per_sm, int iters, double *ms);

/* Quadratic forms q[i] = r_i^T P r_i of n HOST rows R [n][64] against one HOST matrix P [64][64] (p = 60 padded with
 * zeros): kind 0 = warp-tiled FMA (k_chi2's arrangement), kind 1 = FP64 DMMA (mma.sync m8n8k4).  ms = average kernel
 * time over `reps` launches; q = results (for the equality check). */
int vb200p_quad_probe(int device, int kind, const double *R, const double *P, int64_t n, int reps, double *q, double *ms);

/* FP64 FMA issue-rate probe (8 independent DFMA chains per thread, whole GPU).  tflops counts 2 flop
 * per FMA.  A side figure: rooflines are quoted against the nominal FP64 peak. */
int vb200p_fp64_peak(int device, int iters, double *tflops, double *ms);

#ifdef __cplusplus
}
#endif
#endif /* VICTOR_B200_PROBES_H */
