// k1_wide.cu: kernel instantiations of one family (see kernels.h); compiled as its own translation unit.
#include "k1_streaming.cuh"
#include "kernels.h"

namespace vb200 {

// Tuned kernel, the other velocity-integral setups on model coordinates: anisotropic streaming
// (xi_0 + xi_2 L_2 [+ xi_4 L_4]) and the dispersion model.  Fast math only (the libm test variant of these
// models is the general kernel).
// (the dispersion model keeps the cubic refinement whatever the streaming default is: its coordinate iteration
// amplifies seed errors, see k1_streaming.cuh: disp_nodes)
// Nodes per trip / register budget of the two BOSS-shaped instantiations as measured (profiles/r02za_variants_ilp_blocks.txt);
// tables with bucket flags (knots inside buckets: the comparison scan needs registers) keep four nodes at 64 registers.
template <bool kFlags>
k1_fn k1_wide_variant(int rsd_model, int n_ell) {
    if (rsd_model == kRsdDispersion) {
        if (n_ell == 1)   // eight nodes at 128 registers: 25.47 vs 25.90 ms per 16,384 rows
            return k_multipoles<K1Cfg<true, kFlags, kFlags ? 4 : 8, kDefExp, 3, kRsdDispersion, 1, kFlags ? 4 : 2>>;
        if (n_ell == 2) return k_multipoles<K1Cfg<true, kFlags, 4, kDefExp, 3, kRsdDispersion, 2>>;
        return k_multipoles<K1Cfg<true, kFlags, 4, kDefExp, 3, kRsdDispersion, 3>>;
    }
    if (n_ell == 2)   // ten nodes at 80 registers: 8.47 vs 8.59 ms per 16,384 rows
        return k_multipoles<K1Cfg<true, kFlags, kFlags ? 4 : 10, kDefExp, kDefNewton, kRsdStreaming, 2, kFlags ? 4 : 3>>;
    return k_multipoles<K1Cfg<true, kFlags, kFlags ? 4 : 10, kDefExp, kDefNewton, kRsdStreaming, 3, kFlags ? 4 : 3>>;
}

// real-space ccf measured from data (ccf_model.py:675-679) on the tuned kernel: knots on the bucket lattice only
// (the shipped measured-model files).  Same node counts and register budgets as the model-coordinate kernels
// (profiles/r02t_general_kernel_configs.txt)
k1_fn k1_fromdata_variant(int rsd_model, int n_ell) {
    if (rsd_model == kRsdDispersion) {
        if (n_ell == 1) return k_multipoles<K1Cfg<true, false, 8, kDefExp, 3, kRsdDispersion, 1, 2, true>>;
        if (n_ell == 2) return k_multipoles<K1Cfg<true, false, 8, kDefExp, 3, kRsdDispersion, 2, 2, true>>;
        return k_multipoles<K1Cfg<true, false, 2, kDefExp, 3, kRsdDispersion, 3, 4, true>>;
    }
    if (n_ell == 1) return k_multipoles<K1Cfg<true, false, 10, kDefExp, kDefNewton, kRsdStreaming, 1, 3, true>>;
    if (n_ell == 2) return k_multipoles<K1Cfg<true, false, 10, kDefExp, kDefNewton, kRsdStreaming, 2, 3, true>>;
    return k_multipoles<K1Cfg<true, false, 10, kDefExp, kDefNewton, kRsdStreaming, 3, 3, true>>;
}

k1_fn pick_k1_wide(int rsd_model, int n_ell, bool flags, bool from_data) {
    if (from_data) return k1_fromdata_variant(rsd_model, n_ell);
    return flags ? k1_wide_variant<true>(rsd_model, n_ell) : k1_wide_variant<false>(rsd_model, n_ell);
}

}  // namespace vb200
