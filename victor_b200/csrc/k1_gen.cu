// k1_gen.cu: kernel instantiations of one family (see kernels.h); compiled as its own translation unit.
#include "k1_general.cuh"
#include "kernels.h"

namespace vb200 {

k1_fn pick_general_fused(int rsd_model, bool fast) {
    if (!fast) return nullptr;
    if (rsd_model == kRsdStreaming) return k_multipoles_general<kRsdStreaming, true, true>;
    if (rsd_model == kRsdDispersion) return k_multipoles_general<kRsdDispersion, true, true>;
    return k_multipoles_kaiser<true, true>;
}

k1_fn pick_general(int rsd_model, bool fast) {
    if (rsd_model == kRsdStreaming)
        return fast ? k_multipoles_general<kRsdStreaming, true> : k_multipoles_general<kRsdStreaming, false>;
    if (rsd_model == kRsdDispersion)
        return fast ? k_multipoles_general<kRsdDispersion, true> : k_multipoles_general<kRsdDispersion, false>;
    return fast ? k_multipoles_kaiser<true> : k_multipoles_kaiser<false>;
}

}  // namespace vb200
