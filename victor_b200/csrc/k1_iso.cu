// k1_iso.cu: kernel instantiations of one family (see kernels.h); compiled as its own translation unit.
#include "k1_streaming.cuh"
#include "kernels.h"

namespace vb200 {

// Nodes in flight and register budget of the default streaming kernel (profiles/r02za_variants_ilp_blocks.txt):
// ten nodes per trip (nx = 50: five trips, no tail), 80 registers / three 256-thread blocks' worth, launched as
// 128-thread blocks: 27.65 ms per 65,536 rows against 28.15 ms for four nodes at 64 registers.
constexpr int kIsoU = 10, kIsoMinBlocks = 3;

template <bool kFlags>
K1Pick k1_iso_variant(bool fast, int ilp, int expdeg, int newton) {
    if (!fast) return {k_multipoles<K1Cfg<false, kFlags, 1, 6>>, 6};
    if constexpr (kFlags) {   // tables with knots inside buckets: the four-node kernel only
        return {k_multipoles<K1Cfg<true, kFlags, 4, kDefExp, kDefNewton>>, kDefExp};
    } else {
        if (ilp >= 1 && ilp < 4) return {k_multipoles<K1Cfg<true, kFlags, 1, kDefExp, kDefNewton>>, kDefExp};
        if (ilp >= 4 && ilp < kIsoU && expdeg == kDefExp && newton == kDefNewton)   // the four-node kernel (first default of round 2)
            return {k_multipoles<K1Cfg<true, kFlags, 4, kDefExp, kDefNewton>>, kDefExp};
#define VB_V(E, N) \
    if (expdeg == E && newton == N) return {k_multipoles<K1Cfg<true, kFlags, kIsoU, E, N, kRsdStreaming, 1, kIsoMinBlocks>>, E};
        VB_V(5, 3) VB_V(5, 2) VB_V(3, 2)
#undef VB_V
        return {k_multipoles<K1Cfg<true, kFlags, kIsoU, kDefExp, kDefNewton, kRsdStreaming, 1, kIsoMinBlocks>>, kDefExp};
    }
}

K1Pick pick_k1(bool fast, bool flags, int ilp, int expdeg, int newton) {
    if (!expdeg) expdeg = kDefExp;
    if (!newton) newton = kDefNewton;
    return flags ? k1_iso_variant<true>(fast, ilp, expdeg, newton) : k1_iso_variant<false>(fast, ilp, expdeg, newton);
}

// fused likelihood epilogue: for the default tuned configuration only (nullptr otherwise: the caller then
// launches K2 after the plain kernel)
k1_fn pick_k1_fused(bool fast, bool flags, int ilp, int expdeg, int newton) {
    if (!expdeg) expdeg = kDefExp;
    if (!newton) newton = kDefNewton;
    if (!fast || (ilp >= 1 && ilp < 4) || expdeg != kDefExp || newton != kDefNewton) return nullptr;
    return flags ? k_multipoles<K1Cfg<true, true, 4, kDefExp, kDefNewton>, true>
                 : k_multipoles<K1Cfg<true, false, 4, kDefExp, kDefNewton>, true>;
}

}  // namespace vb200
