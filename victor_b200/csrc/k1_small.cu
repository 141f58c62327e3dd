// k1_small.cu: kernel instantiations of one family (see kernels.h); compiled as its own translation unit.
#include "k1_small.cuh"
#include "kernels.h"

namespace vb200 {

// k_small: the few-rows kernel, default math of the tuned families (fast only)
template <bool kFlags>
small_fn small_variant(int rsd_model, int n_ell) {
    if (rsd_model == kRsdDispersion) {
        if (n_ell == 1) return k_small<K1Cfg<true, kFlags, 4, kDefExp, 3, kRsdDispersion, 1>>;
        if (n_ell == 2) return k_small<K1Cfg<true, kFlags, 4, kDefExp, 3, kRsdDispersion, 2>>;
        return k_small<K1Cfg<true, kFlags, 4, kDefExp, 3, kRsdDispersion, 3>>;
    }
    if (n_ell == 1) return k_small<K1Cfg<true, kFlags, 4, kDefExp, kDefNewton>>;
    if (n_ell == 2) return k_small<K1Cfg<true, kFlags, 4, kDefExp, kDefNewton, kRsdStreaming, 2>>;
    return k_small<K1Cfg<true, kFlags, 4, kDefExp, kDefNewton, kRsdStreaming, 3>>;
}
small_fn small_fromdata_variant(int rsd_model, int n_ell) {
    if (rsd_model == kRsdDispersion) {
        if (n_ell == 1) return k_small<K1Cfg<true, false, 4, kDefExp, 3, kRsdDispersion, 1, 4, true>>;
        if (n_ell == 2) return k_small<K1Cfg<true, false, 4, kDefExp, 3, kRsdDispersion, 2, 4, true>>;
        return k_small<K1Cfg<true, false, 4, kDefExp, 3, kRsdDispersion, 3, 4, true>>;
    }
    if (n_ell == 1) return k_small<K1Cfg<true, false, 4, kDefExp, kDefNewton, kRsdStreaming, 1, 4, true>>;
    if (n_ell == 2) return k_small<K1Cfg<true, false, 4, kDefExp, kDefNewton, kRsdStreaming, 2, 4, true>>;
    return k_small<K1Cfg<true, false, 4, kDefExp, kDefNewton, kRsdStreaming, 3, 4, true>>;
}
small_fn pick_small(int rsd_model, int n_ell, bool flags, bool from_data) {
    if (from_data) return small_fromdata_variant(rsd_model, n_ell);   // (lattice knot sets only: see kernel_family)
    return flags ? small_variant<true>(rsd_model, n_ell) : small_variant<false>(rsd_model, n_ell);
}

}  // namespace vb200
