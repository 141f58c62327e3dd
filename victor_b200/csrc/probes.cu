// Measurement and self-test entry points (include/victor_b200_probes.h): NOT part of the product ABI.
// Built into its own library, libvictor_b200_probes.so, from the same device math as the kernels
// (common.cuh) so that tests can hold the hand-rolled exp / rsqrt / rcp to a few ulp and tools can
// measure the FP64 issue model the kernel schedule was designed against.
#include <cmath>
#include <string>

#include "../../include/victor_b200_probes.h"
#include "probes.cuh"

using namespace vb200;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

constexpr int kOk = 0, kEInval = -1, kECuda = -2;

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return fail(kECuda, std::string(#call) + ": " + cudaGetErrorString(e__));     \
    } while (0)

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

// device buffer / event pair released on every return path
struct DevBuf {
    double *p = nullptr;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t doubles) { return cudaMalloc(&p, doubles * sizeof(double)); }
};
struct EventPair {
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() {
        if (a) cudaEventDestroy(a);
        if (b) cudaEventDestroy(b);
    }
    cudaError_t create() {
        cudaError_t e = cudaEventCreate(&a);
        return e != cudaSuccess ? e : cudaEventCreate(&b);
    }
};

void fill_exp_tables(double *t) {   // 2^(j/32) then 2^(j/1024)
    for (int j = 0; j < kExpTab; ++j) t[j] = std::exp2((double)j / kExpTab);
    for (int j = 0; j < kExpTabBig; ++j) t[kExpTab + j] = std::exp2((double)j / kExpTabBig);
}

}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char *vb200p_last_error(void) { return g_err.c_str(); }

int vb200p_math_selftest(int device, const double *x, int64_t n, double *out) {
    if (!x || !out || n <= 0) return fail(kEInval, "bad arguments");
    DeviceGuard g(device);
    if (!g.ok) return fail(kECuda, "cudaSetDevice failed");
    DevBuf dx, dout, dtab;
    static double etab[kExpTab + kExpTabBig];
    fill_exp_tables(etab);
    CK(dx.alloc(n));
    CK(dout.alloc(kSelftestOutputs * n));
    CK(dtab.alloc(kExpTab + kExpTabBig));
    CK(cudaMemcpy(dx.p, x, n * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dtab.p, etab, sizeof(etab), cudaMemcpyHostToDevice));
    k_math_selftest<<<(unsigned)((n + 255) / 256), 256>>>(dx.p, n, dtab.p, dout.p);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dout.p, kSelftestOutputs * n * sizeof(double), cudaMemcpyDeviceToHost));
    return kOk;
}

int vb200p_pipe_probe(int device, int mode, int iters, double *ms) {
    if (!ms || iters < 1 || mode < 0 || mode > 6) return fail(kEInval, "bad arguments");
    DeviceGuard g(device);
    if (!g.ok) return fail(kECuda, "cudaSetDevice failed");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    DevBuf buf;
    CK(buf.alloc(1));
    double *d = buf.p;
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    EventPair ev;
    CK(ev.create());
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    auto run = [&](int it) {
        switch (mode) {
            case 0: k_pipe_probe<0><<<blocks, threads>>>(d, it, 0.999999, 1e-9); break;
            case 1: k_pipe_probe<1><<<blocks, threads>>>(d, it, 0.999999, 1e-9); break;
            case 2: k_pipe_probe<2><<<blocks, threads>>>(d, it, 0.999999, 1e-9); break;
            case 3: k_pipe_probe<3><<<blocks, threads>>>(d, it, 0.999999, 1e-9); break;
            case 4: k_pipe_probe<4><<<blocks, threads>>>(d, it, 0.999999, 1e-9); break;
            case 5: k_pipe_probe<5><<<blocks, threads>>>(d, it, 0.999999, 1e-9); break;
            default: k_pipe_probe<6><<<blocks, threads>>>(d, it, 0.999999, 1e-9); break;
        }
    };
    run(iters / 4 + 1);
    CK(cudaEventRecord(e0));
    run(iters);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, e0, e1));
    *ms = t;
    return kOk;
}

int vb200p_mix_probe(int device, int chains, int mix, int kind, int blocks_per_sm, int iters, double *ms) {
    if (!ms || iters < 1 || blocks_per_sm < 1) return fail(kEInval, "bad arguments");
    DeviceGuard g(device);
    if (!g.ok) return fail(kECuda, "cudaSetDevice failed");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    DevBuf buf;
    CK(buf.alloc(128));
    double *d = buf.p;
    {
        double h[128];
        for (int i = 0; i < 128; ++i) h[i] = (i < 33) ? 0.999999 - 1e-9 * i : 1e-9 + 1e-12 * i;
        CK(cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice));
    }
    const int blocks = prop.multiProcessorCount * blocks_per_sm;
    typedef void (*fn_t)(double *, int, double, double, int);
    fn_t fn = nullptr;
#define VB_MIX(C, M, K) if (chains == C && mix == M && kind == K) fn = k_mix_probe<C, M, K>;
    VB_MIX(1, 0, 0) VB_MIX(2, 0, 0) VB_MIX(4, 0, 0) VB_MIX(8, 0, 0)
    VB_MIX(4, 1, 0) VB_MIX(4, 2, 0) VB_MIX(8, 1, 0) VB_MIX(2, 1, 0)
    VB_MIX(4, 1, 1) VB_MIX(8, 1, 1) VB_MIX(2, 1, 1)
    VB_MIX(4, 0, 2) VB_MIX(8, 0, 2)
    VB_MIX(4, 0, 3) VB_MIX(4, 0, 4)
    VB_MIX(8, 0, 5) VB_MIX(8, 0, 6) VB_MIX(8, 0, 7) VB_MIX(4, 0, 5) VB_MIX(4, 0, 6) VB_MIX(4, 0, 7)
#undef VB_MIX
    if (!fn) return fail(kEInval, "no such probe variant");
    EventPair ev;
    CK(ev.create());
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    fn<<<blocks, 128>>>(d, iters / 4 + 1, 0.999999, 1e-9, 3);
    CK(cudaEventRecord(e0));
    fn<<<blocks, 128>>>(d, iters, 0.999999, 1e-9, 3);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, e0, e1));
    *ms = t;
    return kOk;
}

int vb200p_load_probe(int device, int path, int distinct, int stride, int blocks_per_sm, int iters, double *ms) {
    if (!ms || iters < 1 || distinct < 1 || distinct > 32 || blocks_per_sm < 1 || (path != 0 && path != 1))
        return fail(kEInval, "bad arguments");
    DeviceGuard g(device);
    if (!g.ok) return fail(kECuda, "cudaSetDevice failed");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    DevBuf buf;
    CK(buf.alloc(1));
    LoadProbeArgs a{};
    for (int i = 0; i < 128 * 4; ++i) a.tab[i] = 1.0 / (1.0 + i);
    a.out = buf.p;
    a.distinct = distinct;
    a.stride = stride;
    const int blocks = prop.multiProcessorCount * blocks_per_sm;
    EventPair ev;
    CK(ev.create());
    auto run = [&](int it) {
        a.iters = it;
        if (path == 0) k_load_probe<0><<<blocks, 256>>>(a);
        else k_load_probe<1><<<blocks, 256>>>(a);
    };
    run(iters / 4 + 1);
    CK(cudaEventRecord(ev.a));
    run(iters);
    CK(cudaEventRecord(ev.b));
    CK(cudaEventSynchronize(ev.b));
    CK(cudaGetLastError());
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, ev.a, ev.b));
    *ms = t;
    return kOk;
}

int vb200p_quad_probe(int device, int kind, const double *R, const double *P, int64_t n, int reps, double *q, double *ms) {
    if (!R || !P || !q || !ms || n < 1 || reps < 1 || (kind != 0 && kind != 1)) return fail(kEInval, "bad arguments");
    DeviceGuard g(device);
    if (!g.ok) return fail(kECuda, "cudaSetDevice failed");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    DevBuf dR, dP, dq;
    CK(dR.alloc((size_t)n * kQP));
    CK(dP.alloc((size_t)kQP * kQP));
    CK(dq.alloc((size_t)n));
    CK(cudaMemcpy(dR.p, R, (size_t)n * kQP * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dP.p, P, (size_t)kQP * kQP * sizeof(double), cudaMemcpyHostToDevice));
    const int blocks = prop.multiProcessorCount * 4;
    EventPair ev;
    CK(ev.create());
    auto run = [&]() {
        if (kind == 0) k_quad_fma<<<blocks, 256>>>(dR.p, dP.p, n, dq.p);
        else k_quad_dmma<<<blocks, 256>>>(dR.p, dP.p, n, dq.p);
    };
    run();
    CK(cudaEventRecord(ev.a));
    for (int i = 0; i < reps; ++i) run();
    CK(cudaEventRecord(ev.b));
    CK(cudaEventSynchronize(ev.b));
    CK(cudaGetLastError());
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, ev.a, ev.b));
    *ms = t / reps;
    CK(cudaMemcpy(q, dq.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    return kOk;
}

int vb200p_seed_probe(int device, const double *x, int64_t n, double *out) {
    if (!x || !out || n <= 0) return fail(kEInval, "bad arguments");
    DeviceGuard g(device);
    if (!g.ok) return fail(kECuda, "cudaSetDevice failed");
    DevBuf dx, dout;
    CK(dx.alloc(n));
    CK(dout.alloc(2 * n));
    CK(cudaMemcpy(dx.p, x, n * sizeof(double), cudaMemcpyHostToDevice));
    k_seed_probe<<<(unsigned)((n + 255) / 256), 256>>>(dx.p, n, dout.p);
    CK(cudaGetLastError());
    CK(cudaMemcpy(out, dout.p, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
    return kOk;
}

int vb200p_fp64_peak(int device, int iters, double *tflops, double *ms) {
    if (!tflops || iters < 1) return fail(kEInval, "bad arguments");
    DeviceGuard g(device);
    if (!g.ok) return fail(kECuda, "cudaSetDevice failed");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    DevBuf buf;
    CK(buf.alloc(1));
    double *d = buf.p;
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    EventPair ev;
    CK(ev.create());
    cudaEvent_t e0 = ev.a, e1 = ev.b;
    k_fp64_peak<<<blocks, threads>>>(d, iters / 4 + 1, 0.999999, 1e-9);  // warm-up
    CK(cudaEventRecord(e0));
    k_fp64_peak<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, e0, e1));
    const double fma_count = (double)blocks * threads * (double)iters * 64.0;
    *tflops = 2.0 * fma_count / (t * 1e-3) / 1e12;
    if (ms) *ms = t;
    return kOk;
}

}  // extern "C"
#pragma GCC visibility pop
