// C-ABI housekeeping + small layout/metric kernels of libppde_b200.so (see include/ppde_b200.h).
#include "common.cuh"
#include "../../include/ppde_b200.h"
#include "launch.cuh"
#include <thread>
#include <vector>

namespace ppde {
int g_launch_count = 0;

// one-hot float [n,L,20] -> residue index (argmax, first maximum: onehot2seq, data_utils.py:167-175)
__global__ void onehot_to_aa_kernel(const float* __restrict__ x, int n, int L, uint8_t* __restrict__ aa, int stride) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * L) return;
    const int b = (int)(e / L), i = (int)(e - (int64_t)b * L);
    const float* r = x + e * PPDE_Q;
    int best = 0; float bv = r[0];
#pragma unroll
    for (int a = 1; a < PPDE_Q; ++a) if (r[a] > bv) { bv = r[a]; best = a; }
    aa[(int64_t)b * stride + i] = (uint8_t)best;
}

// residue index -> one-hot float [n,L,20] (seqs_to_onehot, data_utils.py:150-157)
__global__ void aa_to_onehot_kernel(const uint8_t* __restrict__ aa, int stride, int n, int L, float* __restrict__ x) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)n * L * PPDE_Q) return;
    const int64_t pos = e / PPDE_Q;
    const int a = (int)(e - pos * PPDE_Q);
    const int b = (int)(pos / L), i = (int)(pos - (int64_t)b * L);
    x[e] = (aa[(int64_t)b * stride + i] == a) ? 1.f : 0.f;
}

// per-chain edit distance to WT (mut_distance, utils.py:5-14) and a 64-bit FNV-1a sequence hash
// (diversity = number of distinct sequences, make_figures.py:38-49). One warp per chain.
__global__ void population_metrics_kernel(const uint8_t* __restrict__ aa, int stride, int n, int L,
                                          const uint8_t* __restrict__ wt, int32_t* __restrict__ dist,
                                          unsigned long long* __restrict__ hash) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n) return;
    const uint8_t* a = aa + (int64_t)b * stride;
    int d = 0;
    for (int i = lane; i < L; i += 32) d += (a[i] != wt[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if (lane == 0) {
        if (dist) dist[b] = d;
        if (hash) {
            unsigned long long h = 1469598103934665603ull;
            for (int i = 0; i < L; ++i) { h ^= a[i]; h *= 1099511628211ull; }
            hash[b] = h;
        }
    }
}

__global__ void counter_add_kernel(int32_t* t, int32_t inc) { *t += inc; }
}  // namespace ppde

using namespace ppde;

extern "C" const char* ppde_version(void) { return "ppde_b200 0.1 (sm_100a)"; }
extern "C" int ppde_last_launch_count(void) { return g_launch_count; }

extern "C" int ppde_onehot_to_aa(const float* x, int32_t n, int32_t L, uint8_t* aa, int32_t aa_stride, void* stream) {
    const int64_t tot = (int64_t)n * L;
    if (tot <= 0) return 0;
    onehot_to_aa_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, L, aa, aa_stride);
    return launch_done();
}

extern "C" int ppde_aa_to_onehot(const uint8_t* aa, int32_t aa_stride, int32_t n, int32_t L, float* x, void* stream) {
    const int64_t tot = (int64_t)n * L * PPDE_Q;
    if (tot <= 0) return 0;
    aa_to_onehot_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(aa, aa_stride, n, L, x);
    return launch_done();
}

extern "C" int ppde_population_metrics(const uint8_t* aa, int32_t aa_stride, int32_t n, int32_t L, const uint8_t* wt,
                                       int32_t* dist, unsigned long long* hash, void* stream) {
    if (n <= 0) return 0;
    population_metrics_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(aa, aa_stride, n, L, wt, dist, hash);
    return launch_done();
}

extern "C" int ppde_counter_add(int32_t* t_dev, int32_t inc, void* stream) {
    counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(t_dev, inc);
    return launch_done();
}

// ---- host side of the boundary: the reference API hands over a float one-hot [n, L, 20] (80 bytes per residue).  When that
// buffer lives in HOST memory it is reduced to residue indices (1 byte per residue) by the host cores BEFORE the copy, and the
// result is expanded after the copy back: 80x fewer bytes cross PCIe (seqs_to_onehot / onehot2seq,
// ppde/third_party/hsu/data_utils.py:150-175, first maximum as torch.argmax).
static void run_threads(int64_t n, int nthreads, void (*fn)(int64_t, int64_t, void*), void* ctx) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    if (n < 4096 || nthreads == 1) { fn(0, n, ctx); return; }
    std::vector<std::thread> th;
    const int64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        const int64_t lo = (int64_t)t * per, hi = lo + per < n ? lo + per : n;
        if (lo < hi) th.emplace_back(fn, lo, hi, ctx);
    }
    for (auto& t : th) t.join();
}
struct HostCodec { const float* x; uint8_t* aa; float* xo; const uint8_t* aai; int64_t L, stride; };

extern "C" int ppde_host_onehot_to_aa(const float* x, int64_t n, int32_t L, uint8_t* aa, int64_t aa_stride, int32_t nthreads) {
    if (n <= 0) return 0;
    if (!x || !aa || L <= 0 || aa_stride < L) return (int)cudaErrorInvalidValue;
    HostCodec c{x, aa, nullptr, nullptr, L, aa_stride};
    run_threads(n, nthreads, [](int64_t lo, int64_t hi, void* p) {
        const HostCodec& c = *static_cast<HostCodec*>(p);
        for (int64_t b = lo; b < hi; ++b) {
            const float* r = c.x + b * c.L * PPDE_Q;
            uint8_t* o = c.aa + b * c.stride;
            for (int64_t i = 0; i < c.L; ++i, r += PPDE_Q) {
                int best = 0; float bv = r[0];
                for (int a = 1; a < PPDE_Q; ++a) if (r[a] > bv) { bv = r[a]; best = a; }
                o[i] = (uint8_t)best;
            }
            for (int64_t i = c.L; i < c.stride; ++i) o[i] = 0;
        }
    }, &c);
    return 0;
}

extern "C" int ppde_host_aa_to_onehot(const uint8_t* aa, int64_t aa_stride, int64_t n, int32_t L, float* x, int32_t nthreads) {
    if (n <= 0) return 0;
    if (!x || !aa || L <= 0 || aa_stride < L) return (int)cudaErrorInvalidValue;
    HostCodec c{nullptr, nullptr, x, aa, L, aa_stride};
    run_threads(n, nthreads, [](int64_t lo, int64_t hi, void* p) {
        const HostCodec& c = *static_cast<HostCodec*>(p);
        for (int64_t b = lo; b < hi; ++b) {
            float* r = c.xo + b * c.L * PPDE_Q;
            const uint8_t* a = c.aai + b * c.stride;
            for (int64_t i = 0; i < c.L; ++i, r += PPDE_Q) {
                for (int q = 0; q < PPDE_Q; ++q) r[q] = 0.f;
                r[a[i] < PPDE_Q ? a[i] : 0] = 1.f;
            }
        }
    }, &c);
    return 0;
}
