// C-ABI host side of libvictor_b200.so: context (device copies of the host-built tables),
// staging of host buffers, kernel launches.  See include/victor_b200.h for the contract.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/victor_b200.h"

#include "k1_general.cuh"
#include "k1_small.cuh"
#include "k1_streaming.cuh"
#include "k2_chi2.cuh"
#include "kernels.h"

using namespace vb200;

static_assert(kNPar == VB200_NPAR, "kernel row stride and VB200_NPAR differ");

namespace {

thread_local std::string g_err;
#ifdef VB200_SMALL_TIMING
unsigned long long *g_stamps = nullptr;   // diagnostic build only
#endif
constexpr int64_t kBucketMinRows = 4096;   // K2: from this many rows on, group the rows by covariance bracket first
constexpr int64_t kSmallCall = 256;   // rows: calls up to this size go through pinned staging

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess)                                                               \
            return fail(VB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));    \
    } while (0)

// NVTX range for the life of a call: shows up in Nsight tools, costs nothing without one attached
struct Range {
    explicit Range(const char *name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// growable device scratch buffer
struct Scratch {
    double *ptr = nullptr;
    size_t cap = 0;  // in doubles
    int ensure(size_t n) {
        if (n <= cap) return VB200_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = std::max(n, (size_t)1024);
        cudaError_t e = cudaMalloc(&ptr, want * sizeof(double));
        if (e != cudaSuccess) {
            ptr = nullptr;
            return fail(VB200_ENOMEM, std::string("cudaMalloc scratch: ") + cudaGetErrorString(e));
        }
        cap = want;
        return VB200_OK;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

bool is_device_ptr(const void *p) {
    if (!p) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

}  // namespace

struct vb200_ctx {
    int device = 0;
    int sm_count = 148;
    bool has_fit = false;
    ModelDev md{};
    FitDev fd{};
    // fit grids (device)
    int fit_ns = 0, fit_nmu = 0, fit_L = 0;
    double *fit_s = nullptr, *fit_mu = nullptr, *fit_sqmu = nullptr, *fit_wmu = nullptr;
    std::vector<void *> owned;
    Scratch sc_params, sc_theory, sc_chi2, sc_lnl, sc_xi, sc_mult, sc_grid, sc_bucket;
    // options
    int opt_fast = 1, opt_nsplit = 0;
    int opt_threads = 0;          // 0 = automatic: 128-thread blocks for the tuned kernels, 256 for the general one
    int opt_ilp = 0;              // 0 = the kernel family's default number of nodes per trip; 4 and 1 select the older variants
    int opt_expdeg = 0, opt_newton = 0;   // 0 = the kernel family's default (kDefExp / kDefNewton; dispersion: cubic)
    int opt_fuse = 1;             // batch mode: chi2 / lnL in the epilogue of K1 instead of a K2 launch
    int opt_tuned = 1;            // 0: force the general kernel (A/B checks of the tuned kernels)
    int opt_bucket = 1;           // K2 of large batches: rows grouped by covariance bracket, matrices served from shared memory
    int opt_tiny = 1;             // calls of <= kSmallRows rows: one launch of k_small instead of K1 + K2 (0: off)
    int opt_mapped = 1;           // k_small writes (chi2 | lnL) straight into page-locked host memory and raises a flag
                                  // the host polls, and takes its one or two rows from the kernel parameters: one
                                  // bare launch, no copy nodes, no stream synchronise.  (Reading the rows from mapped
                                  // host memory was 2.4x slower: 210 blocks x tens of uncached PCIe reads,
                                  // profiles/r02l_small_call_latency.txt.)
    double *pin_dev = nullptr;    // device-side address of `pin` (unified addressing), null if not available
    double *tiny_xi = nullptr;    // k_small scratch: xi(s, mu) of kSmallRows rows, and the rows' ticket counters
    unsigned *tiny_tickets = nullptr;
    int family = 0;               // kernel family the tables are eligible for (kTunedIso / kTunedWide / kGeneral)
    // small-call path (MCMC steps): page-locked staging for the rows in and (chi2 | lnL) out
    double *pin = nullptr, *d_small = nullptr, *d_small_theory = nullptr;
    // the small-call sequence (H2D, K1, K2, D2H) as an instantiated CUDA graph, rebuilt when n or an option changes
    cudaGraphExec_t small_exec = nullptr;
    int64_t small_n = -1;
    long long small_launches = 0;
    cudaStream_t cap_stream = nullptr;
    bool graphs_ok = true;
    int opt_graph = 1;
    // host-bound bulk outputs (theory vectors, xi blocks): rows go in chunks, and each chunk's device-to-host copy
    // runs on `copy_stream` while later chunks compute
    std::vector<double> grid_host;   // what sc_grid holds (s | mu | sqrt(1-mu^2) | wmu of the last vb200_theory call)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t chunk_done[16] = {nullptr};
    int opt_chunks = 0;           // 0 = automatic, 1 = never split, k = force k chunks
    long long launches = 0;
    size_t k1_smem_limit = 0;
    double xw[2 * kMaxNx] = {0};  // host copy of x_m | w_m for the kernel-parameter table
    bool has_flags = false;       // some bucket entry carries the interior-knot flag
};

namespace {

template <class T>
int upload(vb200_ctx *c, const T *host, size_t count, const T **dev) {
    if (count == 0 || host == nullptr) return fail(VB200_EINVAL, "null or empty table passed to vb200_create");
    void *p = nullptr;
    CK(cudaMalloc(&p, count * sizeof(T)));
    c->owned.push_back(p);
    CK(cudaMemcpy(p, host, count * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<const T *>(p);
    return VB200_OK;
}

int check_model(const vb200_model_tables *m) {
    if (!m) return fail(VB200_EINVAL, "model tables are NULL");
    if (m->ncell < 2 || m->nbucket < 1 || m->nx < 3 || m->nbeta < 2 || m->nresc < 2)
        return fail(VB200_EINVAL, "model tables: bad sizes");
    if (m->n_ell < 1 || m->n_ell > VB200_MAX_POLES) return fail(VB200_EINVAL, "model tables: bad n_ell");
    if (m->nx > kMaxNx) return fail(VB200_EUNSUPPORTED, "more than 128 velocity nodes");
    if (m->ncell > 32767) return fail(VB200_EUNSUPPORTED, "too many spline cells");
    if (m->rsd_model < VB200_RSD_STREAMING || m->rsd_model > VB200_RSD_EUCLID)
        return fail(VB200_EINVAL, "model tables: unknown rsd_model");
    if (m->niter < 0 || m->niter > 64) return fail(VB200_EINVAL, "model tables: bad niter");
    if (m->sv_ny < 0 || (m->sv_ny > 0 && (!m->sv2d || !m->sv_ybreaks)))
        return fail(VB200_EINVAL, "model tables: bad sigma_v(r, mu) template");
    if (!(m->inv_h > 0.0) || !(m->iaH > 0.0) || !(m->template_sigma8 > 0.0))
        return fail(VB200_EINVAL, "model tables: bad scalars");
    if (m->growth_mode < 0 || m->growth_mode > 2) return fail(VB200_EINVAL, "model tables: unknown growth_mode");
    if (m->growth_mode == 2 && !(m->template_fsigma8 > 0.0))
        return fail(VB200_EINVAL, "model tables: velocity template needs template_fsigma8");
    if ((m->v0b == nullptr) != (m->d0b == nullptr) || (m->v0b && m->vd_beta_dependent))
        return fail(VB200_EINVAL, "model tables: bad empirical-correction tables");
    return VB200_OK;
}

// which kernel family serves this context with the current options
//   kTunedIso : streaming + isotropic xi + model coordinates + sigma_v(r): every variant of pick_k1
//   kTunedWide: anisotropic streaming or dispersion on model coordinates + sigma_v(r), fast math, "tuned" option on
//   kGeneral  : everything else
enum { kGeneral = 0, kTunedIso = 1, kTunedWide = 2 };
int kernel_family(const vb200_ctx *c);

// blocks per parameter row: one once the rows alone fill the GPU a few times over, else the s range is split
// so that a single row (MCMC step) still spreads over the SMs.  (A wave-count cost model choosing among all
// splits was tried: within +-7 % of this rule over n = 1 ... 16 384, profiles/r01s_probe_nsplit.log.)
int pick_nsplit(const vb200_ctx *c, long long n, int ns, bool pairwise) {
    int nsplit = pairwise ? 0 : c->opt_nsplit;
    if (nsplit <= 0) {
        const long long target = (long long)c->sm_count * 6;
        nsplit = (n >= target) ? 1 : (int)std::min<long long>(ns, (target + n - 1) / n);
    }
    return std::max(1, std::min(nsplit, ns));
}

k1_fn fused_variant(const vb200_ctx *c);

int kernel_family(const vb200_ctx *c) {
    if (!c->opt_tuned) return kGeneral;
    if (c->family == kTunedWide && !c->opt_fast) return kGeneral;
    return c->family;
}

int rec_doubles(const vb200_ctx *c) { return k1_rec_doubles(c->md.rsd_model, c->md.n_ell); }

// exp variant the tuned kernel of this context runs with (the measurement variants exist for kTunedIso without
// bucket flags only; everything else runs the build's default)
int tuned_exp(const vb200_ctx *c) {
    if (kernel_family(c) == kTunedIso)
        return pick_k1(c->opt_fast != 0, c->has_flags, c->opt_ilp, c->opt_expdeg, c->opt_newton).exp;
    return kDefExp;
}
bool big_table(int expdeg) { return expdeg == 3 || expdeg == 30; }

size_t k1_smem_for(const vb200_ctx *c, int jper, int nmu, int fitd) {
    return kernel_family(c) != kGeneral
               ? k1_smem_bytes(c->md.ncell, jper, nmu, c->md.nbucket, fitd, rec_doubles(c),
                               c->opt_fast && big_table(tuned_exp(c)) ? kExpTabBig : kExpTab)
               : k1g_smem_bytes(c->md.ncell, jper, nmu, c->md.nbucket, fitd);
}

// opt_fuse: 0 never, 1 where it pays, 2 always.  Measured (profiles/r01n_variants_fuse.log):
//   general kernel, velocity-integral models: +0.5 % and no theory scratch -> fused by default;
//   general kernel, kaiser / euclid_special (3000 points per row, short blocks): the epilogue's L2 reads sit
//     exposed at the end of every block, 4.54 ms instead of 4.00 ms per 65 536 rows -> not fused;
//   tuned kernel: with the epilogue compiled in, ptxas reads x_m / w_m through LDC into vector registers
//     instead of LDCU into uniform ones (+2.5 register reads per node), 31.6 ms instead of 29.7 -> not fused.
k1_fn fused_variant(const vb200_ctx *c) {
    if (!c->opt_fuse || !c->has_fit) return nullptr;
    const bool always = c->opt_fuse >= 2;
    const int fam = kernel_family(c);
    if (fam == kTunedIso)
        return always ? pick_k1_fused(c->opt_fast != 0, c->has_flags, c->opt_ilp, c->opt_expdeg, c->opt_newton) : nullptr;
    if (fam == kTunedWide) return nullptr;   // K2 is < 1 % of these steps: separate launch
    if (c->md.rsd_model >= kRsdKaiser && !always) return nullptr;
    return pick_general_fused(c->md.rsd_model, c->opt_fast != 0);
}

int launch_k1(vb200_ctx *c, const double *d_params, long long n, const double *d_s, int ns,
              const double *d_mu, const double *d_sqmu, const double *d_wmu, int nmu, int L,
              double *d_xi, double *d_mult, cudaStream_t st, bool pairwise = false, double *d_chi2 = nullptr,
              double *d_lnl = nullptr, bool *fused = nullptr) {
    if (fused) *fused = false;
    if (n <= 0) return VB200_OK;
    int nsplit = pick_nsplit(c, n, ns, pairwise);
    int jper = (ns + nsplit - 1) / nsplit;
    if (pairwise) jper = std::min(ns, std::max(jper, 64));   // one pair per thread: keep at least two full warps
    // the likelihood epilogue needs the whole theory vector in one block
    k1_fn fused_fn = ((d_chi2 || d_lnl) && nsplit == 1 && !pairwise) ? fused_variant(c) : nullptr;
    const bool want_fuse = fused_fn != nullptr;
    const int fitd = want_fuse ? fused_fit_doubles(c->fd.p) : 0;
    auto smem_for = [&](int jp) { return k1_smem_for(c, jp, nmu, fitd); };
    const int fam = kernel_family(c);
    // long s grids: split further until a block's xi(s, mu) stage leaves room for 4 blocks per SM
    // (or, failing that, at least fits)
    const size_t want = std::min<size_t>(c->k1_smem_limit, (size_t)56 * 1024);
    while (jper > 1 && smem_for(jper) > want) jper = (jper + 1) / 2;
    nsplit = (ns + jper - 1) / jper;
    const int npairs = jper * nmu;
    const int threads_opt = c->opt_threads > 0 ? c->opt_threads : (fam != kGeneral ? 128 : 256);
    int threads = std::min(threads_opt, ((npairs + 31) / 32) * 32);
    threads = std::max(32, std::min(threads, 256));
    const size_t smem = smem_for(jper);
    if (smem > c->k1_smem_limit)
        return fail(VB200_EUNSUPPORTED, "grids too large for one block's shared memory (reduce len(s) * len(mu))");
    const long long blocks = n * nsplit;
    if (blocks > 2147483647LL) return fail(VB200_EINVAL, "too many blocks in one launch");

    K1Args a{};
    a.m = c->md;
    a.params = d_params;
    a.n = n;
    a.s = d_s;
    a.mu = d_mu;
    a.sqmu = d_sqmu;
    a.wmu = d_wmu;
    a.ns = ns;
    a.nmu = nmu;
    a.L = L;
    a.jper = jper;
    a.nsplit = nsplit;
    a.pairwise = pairwise ? 1 : 0;
    a.has_flags = c->has_flags ? 1 : 0;
    if (want_fuse && nsplit == 1) {   // (a long s grid may have been split further above)
        a.fuse = 1;
        a.f = c->fd;
        a.chi2 = d_chi2;
        a.lnl = d_lnl;
        if (fused) *fused = true;
    }
    a.xi_out = d_xi;
    a.mult_out = d_mult;
    memcpy(a.xw, c->xw, sizeof(a.xw));
    if (fam != kGeneral && c->opt_fast) {   // the fast kernel's reciprocal of SV carries the scale of the exp argument: take it out of the weights
        const double scale = big_table(tuned_exp(c)) ? kGaussScaleBig : kGaussScale;
        for (int i = 0; i < c->md.nx; ++i) a.xw[kMaxNx + i] = c->xw[kMaxNx + i] / scale;
    }
    k1_fn fn = fam == kTunedIso    ? pick_k1(c->opt_fast != 0, c->has_flags, c->opt_ilp, c->opt_expdeg, c->opt_newton).fn
               : fam == kTunedWide ? pick_k1_wide(c->md.rsd_model, c->md.n_ell, c->has_flags, c->md.from_data != 0)
                                   : pick_general(c->md.rsd_model, c->opt_fast != 0);
    if (a.fuse) fn = fused_fn;
    void *kargs[] = {(void *)&a};
    CK(cudaLaunchKernel((const void *)fn, dim3((unsigned)blocks), dim3(threads), kargs, smem, st));
    CK(cudaGetLastError());
    c->launches++;
    return VB200_OK;
}

int launch_k2(vb200_ctx *c, const double *d_params, const double *d_theory, long long n, double *d_chi2,
              double *d_lnl, cudaStream_t st) {
    if (n <= 0) return VB200_OK;
    K2Args a{};
    a.f = c->fd;
    a.params = d_params;
    a.theory = d_theory;
    a.n = n;
    a.chi2 = d_chi2;
    a.lnl = d_lnl;
    // large batches: rows grouped by covariance bracket first (k_chi2_bucketed); "bucket" 0 keeps k_chi2
    const int nb = c->fd.cov_fixed ? 1 : c->fd.nbeta_cov;
    if (c->opt_bucket && n >= kBucketMinRows && n <= 2147483647LL && k2_stages(c->fd.p) && nb <= kK2MaxBins) {
        const size_t words = (size_t)n + ((size_t)n + 3) / 4 + 2 * kK2MaxBins;   // order | lo8 | bins, in 4-byte words
        int rc = c->sc_bucket.ensure((words + 1) / 2);
        if (rc) return rc;
        a.order = reinterpret_cast<int *>(c->sc_bucket.ptr);
        a.bins = reinterpret_cast<unsigned *>(a.order + n);
        a.lo8 = reinterpret_cast<unsigned char *>(a.bins + 2 * kK2MaxBins);
        c->launches += 2;   // count + scatter in front of the chi-square kernel
    }
    CK(launch_k2_kernels(a, n, c->sm_count, st));
    CK(cudaGetLastError());
    c->launches++;
    return VB200_OK;
}

// does a likelihood call of n rows go through k_small?  (tuned kernel families with their default math only)
bool use_small(const vb200_ctx *c, long long n) {
    return c->opt_tiny && c->has_fit && n >= 1 && n <= kSmallRows && c->opt_fast && kernel_family(c) != kGeneral &&
           c->fit_L * ((c->fit_nmu + kSmallPairs - 1) / kSmallPairs) <= c->fit_nmu &&   // the shares fit the scratch row
           (c->opt_ilp == 0 || c->opt_ilp >= 4) && !c->opt_expdeg && !c->opt_newton && c->opt_nsplit <= 0 &&
           small_smem_bytes(c->md.ncell, c->md.nbucket, c->fd.p, rec_doubles(c), big_table(kDefExp) ? kExpTabBig : kExpTab) <=
               c->k1_smem_limit;
}

// theory (optional), chi2 and lnL of n <= kSmallRows rows on the fit's grids in ONE launch (k1_small.cuh)
// scratch of k_small: allocated outside any stream capture (cudaMalloc is not allowed inside one)
int ensure_small_scratch(vb200_ctx *c) {
    if (!c->tiny_xi) {
        double *xi = nullptr;
        unsigned *tk = nullptr;
        cudaError_t e = cudaMalloc(&xi, (size_t)kSmallRows * c->fit_ns * c->fit_nmu * sizeof(double));
        if (e == cudaSuccess) e = cudaMalloc(&tk, kSmallRows * sizeof(unsigned));
        if (e == cudaSuccess) e = cudaMemset(tk, 0, kSmallRows * sizeof(unsigned));
        if (e != cudaSuccess) {
            if (xi) cudaFree(xi);
            if (tk) cudaFree(tk);
            return fail(VB200_ENOMEM, std::string("k_small scratch: ") + cudaGetErrorString(e));
        }
        c->tiny_xi = xi;
        c->tiny_tickets = tk;
    }
    return VB200_OK;
}

int launch_small(vb200_ctx *c, const double *d_params, long long n, double *d_theory, double *d_chi2, double *d_lnl,
                 cudaStream_t st, unsigned *done = nullptr, const double *inline_rows = nullptr) {
    if (!c->tiny_xi) return fail(VB200_ECUDA, "internal: k_small scratch not allocated");
    K1Args a{};
    a.m = c->md;
    a.params = d_params;
    a.n = n;
    a.s = c->fit_s;
    a.mu = c->fit_mu;
    a.sqmu = c->fit_sqmu;
    a.wmu = c->fit_wmu;
    a.ns = c->fit_ns;
    a.nmu = c->fit_nmu;
    a.L = c->fit_L;
    a.jper = 1;
    a.nsplit = c->fit_ns;
    a.has_flags = c->has_flags ? 1 : 0;
    a.fuse = 1;
    a.f = c->fd;
    a.chi2 = d_chi2;
    a.lnl = d_lnl;
    const bool big = big_table(kDefExp);
    const double scale = big ? kGaussScaleBig : kGaussScale;
    for (int i = 0; i < c->md.nx; ++i) {
        a.xw[i] = c->xw[i];
        a.xw[kMaxNx + i] = c->xw[kMaxNx + i] / scale;
    }
#ifdef VB200_SMALL_TIMING
    if (!g_stamps) cudaMalloc(&g_stamps, 16 * sizeof(unsigned long long));
    SmallArgs sm{g_stamps, c->tiny_xi, c->tiny_tickets, d_theory, done};
#else
    SmallArgs sm{c->tiny_xi, c->tiny_tickets, d_theory, done};
#endif
    sm.inline_rows = inline_rows ? 1 : 0;
    if (inline_rows) memcpy(sm.rows, inline_rows, (size_t)n * VB200_NPAR * sizeof(double));
    const int nchunk = (c->fit_nmu + kSmallPairs - 1) / kSmallPairs;
    const long long blocks = n * c->fit_ns * nchunk;
    const size_t smem = small_smem_bytes(c->md.ncell, c->md.nbucket, c->fd.p, rec_doubles(c), big ? kExpTabBig : kExpTab);
    void *kargs[] = {(void *)&a, (void *)&sm};
    CK(cudaLaunchKernel((const void *)pick_small(c->md.rsd_model, c->md.n_ell, c->has_flags, c->md.from_data != 0), dim3((unsigned)blocks),
                        dim3(kSmallPairs * kSmallLanes), kargs, smem, st));
    c->launches++;
    return VB200_OK;
}

// H2D of the staged rows, K1, K2 (unless fused), D2H of (chi2 | lnL) -- the whole small call on one stream
// page-locked staging area: rows [kSmallCall][NPAR] | chi2, lnL [2][kSmallCall] | flags [kSmallRows] (unsigned)
constexpr size_t kPinDoubles = (size_t)kSmallCall * (VB200_NPAR + 2) + 8;
unsigned *pin_flags(double *pin) { return reinterpret_cast<unsigned *>(pin + (size_t)kSmallCall * (VB200_NPAR + 2)); }

// does this small call run as one bare launch (k_small with the rows in its parameters, results through the mapped
// staging area, completion by flag)?
bool small_is_mapped(const vb200_ctx *c, int64_t n) { return c->opt_mapped && c->pin_dev && use_small(c, n); }

int small_sequence(vb200_ctx *c, int64_t n, cudaStream_t st) {
    int rc;
    double *h_out = c->pin + (size_t)kSmallCall * VB200_NPAR, *d_out = c->d_small + (size_t)kSmallCall * VB200_NPAR;
    CK(cudaMemcpyAsync(c->d_small, c->pin, (size_t)n * VB200_NPAR * sizeof(double), cudaMemcpyHostToDevice, st));
    if (use_small(c, n)) {
        if ((rc = launch_small(c, c->d_small, n, nullptr, d_out, d_out + n, st))) return rc;
    } else {
        bool fused = false;
        if ((rc = launch_k1(c, c->d_small, n, c->fit_s, c->fit_ns, c->fit_mu, c->fit_sqmu, c->fit_wmu, c->fit_nmu,
                            c->fit_L, nullptr, c->d_small_theory, st, false, d_out, d_out + n, &fused)))
            return rc;
        if (!fused && (rc = launch_k2(c, c->d_small, c->d_small_theory, n, d_out, d_out + n, st))) return rc;
    }
    CK(cudaMemcpyAsync(h_out, d_out, (size_t)2 * n * sizeof(double), cudaMemcpyDeviceToHost, st));
    return VB200_OK;
}

void drop_small_graph(vb200_ctx *c) {
    // (a flag-completed small call may still be retiring on its stream: let it finish before its graph goes)
    if (c->small_exec) cudaDeviceSynchronize();
    if (c->small_exec) cudaGraphExecDestroy(c->small_exec);
    c->small_exec = nullptr;
    c->small_n = -1;
}

// capture small_sequence for n rows into an executable graph; on any failure graphs are switched off for
// this context and the caller uses the plain sequence
bool build_small_graph(vb200_ctx *c, int64_t n) {
    drop_small_graph(c);
    if (!c->cap_stream && cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError();
        return c->graphs_ok = false;
    }
    const long long before = c->launches;
    if (cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return c->graphs_ok = false;
    }
    const int rc = small_sequence(c, n, c->cap_stream);
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(c->cap_stream, &graph);
    c->small_launches = c->launches - before;
    c->launches = before;
    if (rc != VB200_OK || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return c->graphs_ok = false;
    }
    const cudaError_t ei = cudaGraphInstantiate(&c->small_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) {
        c->small_exec = nullptr;
        cudaGetLastError();
        return c->graphs_ok = false;
    }
    c->small_n = n;
    return true;
}

// how many row chunks for `bytes` of host-bound output over n rows: ~4 MB per chunk, at most 8, and never fewer
// than 4096 rows per chunk (a chunk boundary idles the GPU for about half a wave of blocks)
int plan_chunks(const vb200_ctx *c, long long n, size_t bytes) {
    if (c->opt_chunks > 0) return (int)std::min<long long>(std::min(c->opt_chunks, 16), std::max<long long>(n, 1));
    long long k = (long long)(bytes / ((size_t)4 << 20));
    k = std::min<long long>(k, n / 4096);
    return (int)std::max<long long>(1, std::min<long long>(k, 8));
}

int ensure_copy_stream(vb200_ctx *c, int nchunks) {
    if (!c->copy_stream) CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < nchunks; ++i)
        if (!c->chunk_done[i]) CK(cudaEventCreateWithFlags(&c->chunk_done[i], cudaEventDisableTiming));
    return VB200_OK;
}

void fill_exp_table(double *t, int n = kExpTab) {
    for (int j = 0; j < n; ++j) t[j] = std::exp2((double)j / n);
}

}  // namespace

// everything else in the library is compiled with hidden visibility: the C ABI below is all it exports
#pragma GCC visibility push(default)
extern "C" {

const char *vb200_version(void) { return "victor_b200 0.1.0 (sm_100a)"; }

const char *vb200_last_error(void) { return g_err.c_str(); }

int vb200_abi_check(int64_t sizeof_model_tables, int64_t sizeof_fit_tables) {
    if (sizeof_model_tables != (int64_t)sizeof(vb200_model_tables) ||
        sizeof_fit_tables != (int64_t)sizeof(vb200_fit_tables))
        return fail(VB200_EINVAL, "table struct size mismatch between binding and library");
    return VB200_OK;
}

int vb200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        g_err = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return 0;
    }
    return n;
}

void vb200_destroy(vb200_ctx *c) {
    if (!c) return;
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    for (void *p : c->owned) cudaFree(p);
    c->sc_params.release();
    c->sc_theory.release();
    c->sc_chi2.release();
    c->sc_lnl.release();
    c->sc_xi.release();
    c->sc_mult.release();
    c->sc_grid.release();
    c->sc_bucket.release();
    if (c->small_exec) cudaGraphExecDestroy(c->small_exec);
    if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (cudaEvent_t e : c->chunk_done)
        if (e) cudaEventDestroy(e);
    if (c->pin) cudaFreeHost(c->pin);
    if (c->d_small) cudaFree(c->d_small);
    if (c->d_small_theory) cudaFree(c->d_small_theory);
    if (c->tiny_xi) cudaFree(c->tiny_xi);
    if (c->tiny_tickets) cudaFree(c->tiny_tickets);
    delete c;
}

int vb200_create(const vb200_model_tables *m, const vb200_fit_tables *f, int device, vb200_ctx **out) {
    if (!out) return fail(VB200_EINVAL, "out is NULL");
    *out = nullptr;
    int rc = check_model(m);
    if (rc) return rc;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(VB200_EINVAL, "no such CUDA device");
    DeviceGuard g(device);
    if (!g.ok) return fail(VB200_ECUDA, "cudaSetDevice failed");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(VB200_EUNSUPPORTED, "this library is built for sm_100a (B200) only");

    vb200_ctx *c = new vb200_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    auto bail = [&](int code) {
        vb200_destroy(c);
        return code;
    };

    ModelDev &d = c->md;
    d.iaH = m->iaH;
    d.s8t = m->template_sigma8;
    d.beta_fixed = m->beta_fixed;
    d.inv_h = m->inv_h;
    d.vel_indep_AP = m->vel_indep_AP;
    d.rsd_model = m->rsd_model;
    d.n_ell = m->n_ell;
    d.beta_dependent = m->beta_dependent;
    d.ncell = m->ncell;
    d.nbucket = m->nbucket;
    d.maxscan = m->maxscan;
    d.nbeta = m->nbeta;
    d.nx = m->nx;
    d.nresc = m->nresc;
    d.from_data = m->realspace_from_data;
    d.kaiser_approx = m->kaiser_approximation;
    d.kaiser_shift = m->kaiser_coord_shift;
    d.niter = m->niter;
    for (int i = 0; i < kMaxPoles; ++i) d.ells[i] = m->ells[i];
    // (beta-dependent or empirically corrected velocity tables only change the prologue: still the tuned kernels)
    {
        const bool vel_integral = m->rsd_model == VB200_RSD_STREAMING || m->rsd_model == VB200_RSD_DISPERSION;
        bool ells_ok = m->ells[0] == 0;
        for (int i = 1; i < m->n_ell; ++i) ells_ok = ells_ok && m->ells[i] == 2 * i;
        c->family = kGeneral;
        if (vel_integral && ells_ok && m->sv_ny == 0) {
            if (!m->realspace_from_data)
                c->family = (m->rsd_model == VB200_RSD_STREAMING && m->n_ell == 1) ? kTunedIso : kTunedWide;
            else
                c->family = kTunedWide;   // (demoted to the general kernel below if some bucket holds an interior knot)
        }
    }
    const size_t nc4 = (size_t)m->ncell * 4;
    if ((rc = upload(c, m->origin, (size_t)m->ncell, &d.origin))) return bail(rc);
    if ((rc = upload(c, m->upper, (size_t)m->ncell, &d.upper))) return bail(rc);
    if ((rc = upload(c, m->bucket_base, (size_t)m->nbucket, &d.bucket_base))) return bail(rc);
    if ((rc = upload(c, m->beta_grid, (size_t)m->nbeta, &d.beta_grid))) return bail(rc);
    if ((rc = upload(c, m->xi_tab, (size_t)m->n_ell * (m->nbeta - 1) * 4 * nc4, &d.xi_tab))) return bail(rc);
    d.vd_beta_dep = m->vd_beta_dependent;
    d.growth_mode = m->growth_mode;
    d.bias = m->bias;
    d.lin_bias = m->linear_bias;
    d.fs8t = m->template_fsigma8;
    d.growth_scale = m->growth_scale;
    d.v0b = d.d0b = nullptr;
    if (m->v0b) {
        if ((rc = upload(c, m->v0b, nc4, &d.v0b))) return bail(rc);
        if ((rc = upload(c, m->d0b, nc4, &d.d0b))) return bail(rc);
    }
    const size_t nvd = m->vd_beta_dependent ? (size_t)(m->nbeta - 1) * 4 * nc4 : nc4;
    if ((rc = upload(c, m->v0, nvd, &d.v0))) return bail(rc);
    if ((rc = upload(c, m->d0, nvd, &d.d0))) return bail(rc);
    if ((rc = upload(c, m->sv, nc4, &d.sv))) return bail(rc);
    d.sv_ny = m->sv_ny;
    if (m->sv_ny > 0) {
        if ((rc = upload(c, m->sv2d, (size_t)m->ncell * m->sv_ny * 16, &d.sv2d))) return bail(rc);
        if ((rc = upload(c, m->sv_ybreaks, (size_t)m->sv_ny + 1, &d.sv_yb))) return bail(rc);
    }
    if ((rc = upload(c, m->x, (size_t)m->nx, &d.x))) return bail(rc);
    if ((rc = upload(c, m->wx, (size_t)m->nx, &d.wx))) return bail(rc);
    if ((rc = upload(c, m->mu_resc, (size_t)m->nresc, &d.mu_resc))) return bail(rc);
    if ((rc = upload(c, m->w_resc, (size_t)m->nresc, &d.w_resc))) return bail(rc);
    double etab[kExpTab];
    fill_exp_table(etab);
    if ((rc = upload(c, etab, (size_t)kExpTab, &d.exp_tab))) return bail(rc);
    {
        std::vector<double> big(kExpTabBig);
        fill_exp_table(big.data(), kExpTabBig);
        if ((rc = upload(c, big.data(), big.size(), &d.exp_tab_big))) return bail(rc);
    }
    for (int i = 0; i < m->nbucket; ++i) c->has_flags = c->has_flags || (m->bucket_base[i] < 0);
    if (m->realspace_from_data && c->has_flags) c->family = kGeneral;   // from-data tuned variants exist without flags only
    for (int i = 0; i < m->nx; ++i) {
        c->xw[i] = m->x[i];
        c->xw[kMaxNx + i] = m->wx[i];
    }

    if (f) {
        if (f->ns < 1 || f->npoles < 1 || f->npoles > VB200_MAX_POLES || f->nmu < 2 || f->nbeta_ccf < 2 ||
            f->nbeta_cov < 1)
            return bail(fail(VB200_EINVAL, "fit tables: bad sizes"));
        const int p = f->ns * f->npoles;
        if (p > 32 * kK2MaxChunks) return bail(fail(VB200_EUNSUPPORTED, "data vector longer than 256"));
        FitDev &fd = c->fd;
        fd.p = p;
        fd.data_beta_dependent = f->data_beta_dependent;
        fd.nbeta_ccf = f->nbeta_ccf;
        fd.cov_fixed = f->cov_fixed;
        fd.nbeta_cov = f->nbeta_cov;
        fd.like_kind = f->like_kind;
        fd.use_logdet = f->use_logdet;
        fd.like_a = f->like_a;
        fd.like_nm1 = f->like_nm1;
        if ((rc = upload(c, f->beta_ccf, (size_t)f->nbeta_ccf, &fd.beta_ccf))) return bail(rc);
        if ((rc = upload(c, f->data_tab, (size_t)(f->nbeta_ccf - 1) * 4 * p, &fd.data_tab))) return bail(rc);
        if ((rc = upload(c, f->beta_cov, (size_t)f->nbeta_cov, &fd.beta_cov))) return bail(rc);
        if ((rc = upload(c, f->icov, (size_t)f->nbeta_cov * p * p, &fd.icov))) return bail(rc);
        if ((rc = upload(c, f->logdet, (size_t)f->nbeta_cov, &fd.logdet))) return bail(rc);
        if ((rc = upload(c, f->lam, (size_t)f->nbeta_cov * p, &fd.lam))) return bail(rc);
        c->fit_ns = f->ns;
        c->fit_nmu = f->nmu;
        c->fit_L = f->npoles;
        std::vector<double> sq((size_t)f->nmu);
        for (int k = 0; k < f->nmu; ++k) sq[k] = std::sqrt(1.0 - f->mu[k] * f->mu[k]);
        const double *tmp = nullptr;
        if ((rc = upload(c, f->s, (size_t)f->ns, &tmp))) return bail(rc);
        c->fit_s = const_cast<double *>(tmp);
        if ((rc = upload(c, f->mu, (size_t)f->nmu, &tmp))) return bail(rc);
        c->fit_mu = const_cast<double *>(tmp);
        if ((rc = upload(c, sq.data(), (size_t)f->nmu, &tmp))) return bail(rc);
        c->fit_sqmu = const_cast<double *>(tmp);
        if ((rc = upload(c, f->wmu, (size_t)f->npoles * f->nmu, &tmp))) return bail(rc);
        c->fit_wmu = const_cast<double *>(tmp);
        c->has_fit = true;
    }

    // allow the large dynamic shared memory carve-out (dense mu grids stage up to ~200 KB)
    c->k1_smem_limit = std::min<size_t>((size_t)prop.sharedMemPerBlockOptin, (size_t)200 * 1024);
    std::vector<const void *> fns;
    const int exps[] = {5, 3};
    for (int fl = 0; fl < 2; ++fl) {
        fns.push_back((const void *)pick_k1(false, fl, 1, 6, 3).fn);
        fns.push_back((const void *)pick_k1(true, fl, 1, kDefExp, kDefNewton).fn);
        for (int e : exps)
            for (int nw = 2; nw <= 3; ++nw) fns.push_back((const void *)pick_k1(true, fl, 4, e, nw).fn);
        fns.push_back((const void *)pick_k1_fused(true, fl, 4, kDefExp, kDefNewton));
        for (int r = 0; r < 2; ++r)
            for (int l = 1; l <= 3; ++l) {
                if (r || l > 1) fns.push_back((const void *)pick_k1_wide(r, l, fl));
                if (!fl) fns.push_back((const void *)pick_k1_wide(r, l, false, true));
            }
    }
    for (int r = 0; r < 6; ++r) fns.push_back((const void *)pick_general(r >> 1, r & 1));
    for (int r = 0; r < 3; ++r) fns.push_back((const void *)pick_general_fused(r, true));
    for (int fl = 0; fl < 2; ++fl)
        for (int r = 0; r < 2; ++r)
            for (int l = 1; l <= 3; ++l) {
                fns.push_back((const void *)pick_small(r, l, fl));
                if (!fl) fns.push_back((const void *)pick_small(r, l, false, true));
            }
    for (const void *fn : fns) {
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->k1_smem_limit);
        if (e != cudaSuccess)
            return bail(fail(VB200_ECUDA, std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e)));
    }
    *out = c;
    return VB200_OK;
}

int vb200_set_option(vb200_ctx *c, const char *key, int64_t value) {
    if (!c || !key) return fail(VB200_EINVAL, "null argument");
    DeviceGuard g(c->device);
    drop_small_graph(c);   // the captured small-call sequence bakes the kernel variant in
    if (!strcmp(key, "graph")) c->opt_graph = value ? 1 : 0;
    else if (!strcmp(key, "chunks")) c->opt_chunks = (int)std::max<int64_t>(0, std::min<int64_t>(value, 16));
    else if (!strcmp(key, "fast_math")) c->opt_fast = value ? 1 : 0;
    else if (!strcmp(key, "nsplit")) c->opt_nsplit = (int)value;
    else if (!strcmp(key, "fuse")) c->opt_fuse = (int)std::max<int64_t>(0, std::min<int64_t>(value, 2));
    else if (!strcmp(key, "tuned")) c->opt_tuned = value ? 1 : 0;
    else if (!strcmp(key, "tiny")) c->opt_tiny = value ? 1 : 0;
    else if (!strcmp(key, "bucket")) c->opt_bucket = value ? 1 : 0;
    else if (!strcmp(key, "mapped")) c->opt_mapped = value ? 1 : 0;
    else if (!strcmp(key, "newton")) {
        if (value != 0 && value != 2 && value != 3)
            return fail(VB200_EINVAL, "newton must be 0 (default), 2 (one Newton step) or 3 (cubic step)");
        c->opt_newton = (int)value;
    } else if (!strcmp(key, "ilp")) c->opt_ilp = value <= 0 ? 0 : (value >= 10 ? 10 : (value >= 4 ? 4 : 1));
    else if (!strcmp(key, "exp_degree")) {
        if (value != 0 && value != 5 && value != 3)
            return fail(VB200_EINVAL, "exp_degree must be 0 (default), 5 or 3 (1024-entry table, degree-3 remainder)");
        c->opt_expdeg = (int)value;
    } else if (!strcmp(key, "threads")) {
        if (value != 0 && (value < 32 || value > 256 || value % 32))
            return fail(VB200_EINVAL, "threads must be 0 (automatic) or 32..256, multiple of 32");
        c->opt_threads = (int)value;
    } else
        return fail(VB200_EINVAL, std::string("unknown option ") + key);
    return VB200_OK;
}

int vb200_synchronize(vb200_ctx *c) {
    if (!c) return fail(VB200_EINVAL, "ctx is NULL");
    DeviceGuard g(c->device);
    CK(cudaDeviceSynchronize());
    return VB200_OK;
}

int64_t vb200_launch_count(const vb200_ctx *c) { return c ? c->launches : 0; }

#ifdef VB200_SMALL_TIMING
extern "C" int vb200_debug_stamps(unsigned long long *out) {   // diagnostic build only: 16 globaltimer stamps
    if (!g_stamps) return -1;
    return cudaMemcpy(out, g_stamps, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -2;
}
#endif

static int theory_impl(vb200_ctx *c, const double *params, int64_t n, const double *s, int32_t ns, const double *mu,
                       int32_t nmu, const double *wmu, int32_t L, double *xi_out, double *mult_out, void *stream,
                       bool pairwise) {
    if (!c) return fail(VB200_EINVAL, "ctx is NULL");
    if (n < 0 || !params || !s || !mu || ns < 1 || nmu < 1) return fail(VB200_EINVAL, "bad arguments");
    if (mult_out && (!wmu || L < 1 || L > VB200_MAX_POLES)) return fail(VB200_EINVAL, "mult_out needs wmu and 1 <= L <= 3");
    if (!xi_out && !mult_out) return fail(VB200_EINVAL, "no output requested");
    if (n == 0) return VB200_OK;
    Range range("vb200_theory");
    DeviceGuard g(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    bool host_io = false;
    int rc;

    // grids: host arrays, staged per call (s | mu | sqrt(1-mu^2) | wmu)
    const int Lw = mult_out ? L : 0;
    const size_t ng = (size_t)ns + 2 * (size_t)nmu + (size_t)Lw * nmu;
    if ((rc = c->sc_grid.ensure(ng))) return rc;
    {
        std::vector<double> h(ng);
        std::copy(s, s + ns, h.begin());
        std::copy(mu, mu + nmu, h.begin() + ns);
        for (int k = 0; k < nmu; ++k) h[(size_t)ns + nmu + k] = std::sqrt(1.0 - mu[k] * mu[k]);
        if (Lw) std::copy(wmu, wmu + (size_t)Lw * nmu, h.begin() + ns + 2 * (size_t)nmu);
        // repeated calls on the same grids (the usual case) find them on the device already
        if (h.size() != c->grid_host.size() || memcmp(h.data(), c->grid_host.data(), ng * sizeof(double)) != 0) {
            // the copy is stream-ordered; the pageable source is consumed before the call returns
            CK(cudaMemcpyAsync(c->sc_grid.ptr, h.data(), ng * sizeof(double), cudaMemcpyHostToDevice, st));
            CK(cudaStreamSynchronize(st));
            c->grid_host.swap(h);
        }
    }
    const double *d_s = c->sc_grid.ptr, *d_mu = d_s + ns, *d_sq = d_mu + nmu;
    const double *d_w = Lw ? d_sq + nmu : nullptr;
    if (pairwise) nmu = 1;   // the kernels see one mu per s: mu[j] belongs to s[j]

    const double *d_params = params;
    if (!is_device_ptr(params)) {
        host_io = true;
        if ((rc = c->sc_params.ensure((size_t)n * VB200_NPAR))) return rc;
        CK(cudaMemcpyAsync(c->sc_params.ptr, params, (size_t)n * VB200_NPAR * sizeof(double),
                           cudaMemcpyHostToDevice, st));
        d_params = c->sc_params.ptr;
    }
    double *d_xi = xi_out, *d_mult = mult_out;
    const size_t nxi = (size_t)n * nmu * ns, nmult = (size_t)n * Lw * ns;
    if (xi_out && !is_device_ptr(xi_out)) {
        host_io = true;
        if ((rc = c->sc_xi.ensure(nxi))) return rc;
        d_xi = c->sc_xi.ptr;
    }
    if (mult_out && !is_device_ptr(mult_out)) {
        host_io = true;
        if ((rc = c->sc_mult.ensure(nmult))) return rc;
        d_mult = c->sc_mult.ptr;
    }
    const bool xi_to_host = xi_out && d_xi != xi_out, mult_to_host = mult_out && d_mult != mult_out;
    const size_t sx = (size_t)nmu * ns, sm = (size_t)Lw * ns;   // doubles per row of each output
    const int nchunks = plan_chunks(c, n, ((xi_to_host ? nxi : 0) + (mult_to_host ? nmult : 0)) * sizeof(double));
    if (nchunks <= 1) {
        if ((rc = launch_k1(c, d_params, n, d_s, ns, d_mu, d_sq, d_w, nmu, Lw, d_xi, d_mult, st, pairwise))) return rc;
        if (xi_to_host) CK(cudaMemcpyAsync(xi_out, d_xi, nxi * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (mult_to_host) CK(cudaMemcpyAsync(mult_out, d_mult, nmult * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (host_io) CK(cudaStreamSynchronize(st));
        return VB200_OK;
    }
    // all chunks are queued first; the (host-blocking, pageable) copies then trail the kernels chunk by chunk
    if ((rc = ensure_copy_stream(c, nchunks))) return rc;
    const int64_t per = (n + nchunks - 1) / nchunks;
    for (int k = 0; k < nchunks; ++k) {
        const int64_t lo = k * per, cnt = std::min<int64_t>(per, n - lo);
        if (cnt <= 0) break;
        if ((rc = launch_k1(c, d_params + lo * VB200_NPAR, cnt, d_s, ns, d_mu, d_sq, d_w, nmu, Lw,
                            d_xi ? d_xi + lo * sx : nullptr, d_mult ? d_mult + lo * sm : nullptr, st, pairwise)))
            return rc;
        CK(cudaEventRecord(c->chunk_done[k], st));
    }
    for (int k = 0; k < nchunks; ++k) {
        const int64_t lo = k * per, cnt = std::min<int64_t>(per, n - lo);
        if (cnt <= 0) break;
        CK(cudaStreamWaitEvent(c->copy_stream, c->chunk_done[k], 0));
        if (xi_to_host)
            CK(cudaMemcpyAsync(xi_out + lo * sx, d_xi + lo * sx, cnt * sx * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        if (mult_to_host)
            CK(cudaMemcpyAsync(mult_out + lo * sm, d_mult + lo * sm, cnt * sm * sizeof(double), cudaMemcpyDeviceToHost,
                               c->copy_stream));
    }
    CK(cudaStreamSynchronize(c->copy_stream));
    CK(cudaStreamSynchronize(st));
    return VB200_OK;
}

int vb200_theory(vb200_ctx *c, const double *params, int64_t n, const double *s, int32_t ns, const double *mu,
                 int32_t nmu, const double *wmu, int32_t L, double *xi_out, double *mult_out, void *stream) {
    return theory_impl(c, params, n, s, ns, mu, nmu, wmu, L, xi_out, mult_out, stream, false);
}

int vb200_theory_pairs(vb200_ctx *c, const double *params, int64_t n, const double *s, const double *mu,
                       int32_t npairs, double *xi_out, void *stream) {
    if (!xi_out) return fail(VB200_EINVAL, "xi_out is NULL");
    return theory_impl(c, params, n, s, npairs, mu, npairs, nullptr, 0, xi_out, nullptr, stream, true);
}

int vb200_likelihood(vb200_ctx *c, const double *params, int64_t n, double *theory, double *chi2, double *lnlike,
                     void *stream) {
    if (!c) return fail(VB200_EINVAL, "ctx is NULL");
    if (!c->has_fit) return fail(VB200_EINVAL, "context was created without fit tables");
    if (n < 0 || !params) return fail(VB200_EINVAL, "bad arguments");
    if (!theory && !chi2 && !lnlike) return fail(VB200_EINVAL, "no output requested");
    if (n == 0) return VB200_OK;
    Range range("vb200_likelihood");
    DeviceGuard g(c->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int p = c->fd.p;
    bool host_io = false;
    int rc;

    const bool params_on_host = !is_device_ptr(params);
    if (n <= kSmallCall && !theory && chi2 && lnlike && params_on_host && !is_device_ptr(chi2) &&
        !is_device_ptr(lnlike)) {
        // one pinned H2D of the rows, two launches, one pinned D2H of (chi2 | lnL), one sync -- replayed as a
        // CUDA graph (one submission instead of four) while n and the options stay the same
        if (!c->pin || !c->d_small || !c->d_small_theory) {
            // all three or none: a failed allocation must not leave a half-initialised staging set behind
            double *pin = nullptr, *d_small = nullptr, *d_theory = nullptr;
            cudaError_t e = cudaHostAlloc(&pin, kPinDoubles * sizeof(double), cudaHostAllocMapped);
            if (e == cudaSuccess) e = cudaMalloc(&d_small, (size_t)kSmallCall * (VB200_NPAR + 2) * sizeof(double));
            if (e == cudaSuccess) e = cudaMalloc(&d_theory, (size_t)kSmallCall * p * sizeof(double));
            if (e != cudaSuccess) {
                if (pin) cudaFreeHost(pin);
                if (d_small) cudaFree(d_small);
                if (d_theory) cudaFree(d_theory);
                return fail(VB200_ENOMEM, std::string("small-call staging: ") + cudaGetErrorString(e));
            }
            c->pin = pin;
            c->d_small = d_small;
            c->d_small_theory = d_theory;
            void *dp = nullptr;
            if (cudaHostGetDevicePointer(&dp, pin, 0) == cudaSuccess) c->pin_dev = static_cast<double *>(dp);
            else cudaGetLastError();
        }
        if (use_small(c, n) && (rc = ensure_small_scratch(c))) return rc;
        double *h_out = c->pin + (size_t)kSmallCall * VB200_NPAR;
        memcpy(c->pin, params, (size_t)n * VB200_NPAR * sizeof(double));
        const bool mapped = small_is_mapped(c, n);
        volatile unsigned *flags = pin_flags(c->pin);
        if (mapped)
            for (int64_t i = 0; i < n; ++i) flags[i] = 0u;
        bool replayed = false;
        if (mapped) {
            // one launch and nothing else on the stream: the rows ride in the kernel parameters, the results come back
            // through mapped memory (a graph would only wrap this single node)
            double *m_out = c->pin_dev + (size_t)kSmallCall * VB200_NPAR;
            if ((rc = launch_small(c, nullptr, n, nullptr, m_out, m_out + n, st, pin_flags(c->pin_dev), params))) return rc;
            replayed = true;
        } else if (c->opt_graph && c->graphs_ok && (c->small_n == n || build_small_graph(c, n))) {
            if (cudaGraphLaunch(c->small_exec, st) == cudaSuccess) {
                c->launches += c->small_launches;
                replayed = true;
            } else {
                cudaGetLastError();
                c->graphs_ok = false;
            }
        }
        if (!replayed && (rc = small_sequence(c, n, st))) return rc;
        bool done = false;
        if (mapped) {
            // the last block of every row raises its flag after a system-scope fence behind its results: poll the
            // flags (a few microseconds) instead of paying the wake-up latency of a stream synchronise; give up after
            // ~2 ms and synchronise, which also reports a failed launch
            for (long spin = 0; spin < 4000000 && !done; ++spin) {
                done = true;
                for (int64_t i = 0; i < n; ++i) done = done && flags[i] != 0u;
            }
        }
        if (!done) CK(cudaStreamSynchronize(st));
        std::atomic_thread_fence(std::memory_order_acquire);   // results are read after the flags, not before
        memcpy(chi2, h_out, (size_t)n * sizeof(double));
        memcpy(lnlike, h_out + n, (size_t)n * sizeof(double));
        return VB200_OK;
    }

    const double *d_params = params;
    if (params_on_host) {
        host_io = true;
        if ((rc = c->sc_params.ensure((size_t)n * VB200_NPAR))) return rc;
        CK(cudaMemcpyAsync(c->sc_params.ptr, params, (size_t)n * VB200_NPAR * sizeof(double),
                           cudaMemcpyHostToDevice, st));
        d_params = c->sc_params.ptr;
    }
    double *d_theory = theory;
    // chi2 / lnL from the epilogue of K1: the theory vectors only leave the SM if the caller wants them
    const bool will_fuse = (chi2 || lnlike) && pick_nsplit(c, n, c->fit_ns, false) == 1 && fused_variant(c) &&
                           k1_smem_for(c, c->fit_ns, c->fit_nmu, fused_fit_doubles(p)) <=
                               std::min<size_t>(c->k1_smem_limit, (size_t)56 * 1024);
    if (!theory && will_fuse) {
        d_theory = nullptr;
    } else if (!theory || !is_device_ptr(theory)) {
        if ((rc = c->sc_theory.ensure((size_t)n * p))) return rc;
        d_theory = c->sc_theory.ptr;
        if (theory) host_io = true;
    }
    double *d_chi2 = chi2, *d_lnl = lnlike;
    if (chi2 && !is_device_ptr(chi2)) {
        host_io = true;
        if ((rc = c->sc_chi2.ensure((size_t)n))) return rc;
        d_chi2 = c->sc_chi2.ptr;
    }
    if (lnlike && !is_device_ptr(lnlike)) {
        host_io = true;
        if ((rc = c->sc_lnl.ensure((size_t)n))) return rc;
        d_lnl = c->sc_lnl.ptr;
    }
    const bool theory_to_host = theory && d_theory != theory;
    if ((chi2 || lnlike) && use_small(c, n)) {
        // a handful of rows (device buffers, or the theory vectors wanted too): still one launch
        if ((rc = ensure_small_scratch(c))) return rc;
        if ((rc = launch_small(c, d_params, n, theory ? d_theory : nullptr, d_chi2, d_lnl, st))) return rc;
        if (theory_to_host) CK(cudaMemcpyAsync(theory, d_theory, (size_t)n * p * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (chi2 && d_chi2 != chi2) CK(cudaMemcpyAsync(chi2, d_chi2, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (lnlike && d_lnl != lnlike)
            CK(cudaMemcpyAsync(lnlike, d_lnl, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
        if (host_io) CK(cudaStreamSynchronize(st));
        return VB200_OK;
    }
    const int nchunks = theory_to_host ? plan_chunks(c, n, (size_t)n * p * sizeof(double)) : 1;
    if (nchunks > 1 && (rc = ensure_copy_stream(c, nchunks))) return rc;
    const int64_t per = (n + nchunks - 1) / nchunks;
    for (int k = 0; k < nchunks; ++k) {
        const int64_t lo = k * per, cnt = std::min<int64_t>(per, n - lo);
        if (cnt <= 0) break;
        bool fused = false;
        if ((rc = launch_k1(c, d_params + lo * VB200_NPAR, cnt, c->fit_s, c->fit_ns, c->fit_mu, c->fit_sqmu, c->fit_wmu,
                            c->fit_nmu, c->fit_L, nullptr, d_theory ? d_theory + lo * p : nullptr, st, false,
                            d_chi2 ? d_chi2 + lo : nullptr, d_lnl ? d_lnl + lo : nullptr, &fused)))
            return rc;
        if ((chi2 || lnlike) && !fused) {
            if (!d_theory) return fail(VB200_ECUDA, "internal: likelihood epilogue was expected to be fused");
            if ((rc = launch_k2(c, d_params + lo * VB200_NPAR, d_theory + lo * p, cnt, d_chi2 ? d_chi2 + lo : nullptr,
                                d_lnl ? d_lnl + lo : nullptr, st)))
                return rc;
        }
        if (nchunks > 1) CK(cudaEventRecord(c->chunk_done[k], st));
    }
    if (theory_to_host && nchunks > 1) {
        // the copies (host-blocking for pageable memory) trail the kernels chunk by chunk on their own stream
        for (int k = 0; k < nchunks; ++k) {
            const int64_t lo = k * per, cnt = std::min<int64_t>(per, n - lo);
            if (cnt <= 0) break;
            CK(cudaStreamWaitEvent(c->copy_stream, c->chunk_done[k], 0));
            CK(cudaMemcpyAsync(theory + lo * p, d_theory + lo * p, (size_t)cnt * p * sizeof(double), cudaMemcpyDeviceToHost,
                               c->copy_stream));
        }
        CK(cudaStreamSynchronize(c->copy_stream));
    } else if (theory_to_host) {
        CK(cudaMemcpyAsync(theory, d_theory, (size_t)n * p * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    if (chi2 && d_chi2 != chi2) CK(cudaMemcpyAsync(chi2, d_chi2, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (lnlike && d_lnl != lnlike)
        CK(cudaMemcpyAsync(lnlike, d_lnl, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (host_io) CK(cudaStreamSynchronize(st));
    return VB200_OK;
}

}  // extern "C"
#pragma GCC visibility pop
