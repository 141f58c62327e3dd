// Host-side view of the kernel variants of libvictor_b200.so.  The kernels are instantiated in separate translation
// units (k1_iso.cu, k1_wide.cu, k1_small.cu, k1_gen.cu, k2.cu) so that the library builds in parallel; api.cu only sees
// the function pointers handed out here.
#pragma once
#include "common.cuh"

// defaults of the tuned kernel's math (vb200_set_option "exp_degree", "newton" select the others)
#ifndef VB200_DEFAULT_EXP
#define VB200_DEFAULT_EXP 5
#endif
#ifndef VB200_DEFAULT_NEWTON
#define VB200_DEFAULT_NEWTON 2   // one Newton step on the MUFU seeds: +3.9 % and 5.5e-12 / 7.5e-10 measured margins
#endif

namespace vb200 {

struct SmallArgs;
typedef void (*k1_fn)(const K1Args);
typedef void (*small_fn)(const K1Args, const SmallArgs);

constexpr int kDefExp = VB200_DEFAULT_EXP, kDefNewton = VB200_DEFAULT_NEWTON;

// a kernel and the exp variant it was built with (the host folds that variant's argument scale into the weights
// and sizes the exp table in shared memory accordingly)
struct K1Pick {
    k1_fn fn;
    int exp;
};

// k1_iso.cu -- tuned kernel, streaming model + isotropic xi (the BOSS likelihood).  Default <fast, U = 4, exp kDefExp,
// refinement kDefNewton>; the others exist for parity tests (libm math) and for measurement (ILP, exp polynomial,
// refinement order).  Tables with knots off the bucket lattice (flags) get the default and the libm variant only.
K1Pick pick_k1(bool fast, bool flags, int ilp, int expdeg, int newton);
// fused likelihood epilogue: for the default tuned configuration only (nullptr otherwise)
k1_fn pick_k1_fused(bool fast, bool flags, int ilp, int expdeg, int newton);
// k1_wide.cu -- tuned kernel: anisotropic streaming, dispersion, real-space ccf measured from data (fast math only)
k1_fn pick_k1_wide(int rsd_model, int n_ell, bool flags, bool from_data = false);
// k1_small.cu -- the one-launch kernel for calls of one or two rows
small_fn pick_small(int rsd_model, int n_ell, bool flags, bool from_data = false);
// k1_gen.cu -- general kernel, one variant per rsd_model (+ libm, + fused epilogue)
k1_fn pick_general(int rsd_model, bool fast);
k1_fn pick_general_fused(int rsd_model, bool fast);
// k2.cu -- chi-square / likelihood after a K1 launch: one warp per row (batches) or one block per row (few rows)
cudaError_t launch_k2_kernels(const K2Args &a, long long n, int sm_count, cudaStream_t st);

}  // namespace vb200
