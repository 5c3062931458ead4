// Population report of the sampler's log_every path, on the device (one small D2H of the finished numbers).
//
// Reference: PPDE_PAS.run logging block                ppde/protein_samplers/ppde.py:155-170
//              np.quantile(x, [0.5, 0.9]) of energy / predicted fitness / oracle fitness, sum(accepted), mean(dist)
//            n_hops (mean / std edit distance to WT)   scripts/make_figures.py:29-36, ppde/metrics.py:78-85
//            diversity_score (unique sequences / K)    scripts/make_figures.py:38-49
//            top-k of the energies                      ppde/protein_samplers/cmaes.py:34-40 (torch.topk semantics)
// All of it is integer / order-statistic work and exact: radix select on order-preserving keys, an open-addressing table
// that compares whole sequences (a hash only picks the first slot), integer sums.  Under torch.distributed the inputs are the
// all-gathered vectors (ppde_b200/dist.py); every kernel here is single-GPU.
#include "common.cuh"
#include "../../include/ppde_b200.h"
#include "launch.cuh"

namespace ppde {

__device__ __forceinline__ uint32_t f32_key(float v) {           // order-preserving map fp32 -> u32 (-0 == +0, NaN last)
    const uint32_t b = __float_as_uint(v + 0.0f);
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float f32_unkey(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// key of the element with 0-based ascending rank `rank` among x[0..n): MSB-first radix select, 8 bits per pass, one CTA.
// All threads call; hist = 256 ints of shared memory, bc = 2 ints.
template <int NT>
__device__ uint32_t block_select_key(const float* __restrict__ x, int64_t n, int64_t rank, int* hist, int64_t* bc) {
    uint32_t prefix = 0u, pmask = 0u;
    int64_t r = rank;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += NT) hist[i] = 0;
        __syncthreads();
        for (int64_t i = threadIdx.x; i < n; i += NT) {
            const uint32_t k = f32_key(x[i]);
            if ((k & pmask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t acc = 0;
            int b = 0;
            for (; b < 255; ++b) {
                if (acc + hist[b] > r) break;
                acc += hist[b];
            }
            bc[0] = b; bc[1] = r - acc;
        }
        __syncthreads();
        prefix |= (uint32_t)bc[0] << shift;
        pmask |= 255u << shift;
        r = bc[1];
        __syncthreads();
    }
    return prefix;
}

// np.quantile(x, q) with the default 'linear' method, one CTA per quantile:
//   v = q (n - 1), lo = floor(v), t = v - lo, a = x_(lo), b = x_(min(lo + 1, n - 1));  a + (b - a) t,  or  b - (b - a)(1 - t) for
//   t >= 0.5 (numpy's _lerp), evaluated in double like numpy does for a float64 q.
template <int NT>
__global__ void __launch_bounds__(NT) quantile_kernel(const float* __restrict__ x, int64_t n, const double* __restrict__ q, double* __restrict__ out) {
    __shared__ int hist[256];
    __shared__ int64_t bc[2];
    const double v = q[blockIdx.x] * (double)(n - 1);
    int64_t lo = (int64_t)floor(v);
    lo = lo < 0 ? 0 : (lo > n - 1 ? n - 1 : lo);
    const int64_t hi = lo + 1 < n ? lo + 1 : n - 1;
    const double t = v - (double)lo;
    const double a = (double)f32_unkey(block_select_key<NT>(x, n, lo, hist, bc));
    const double b = (hi == lo) ? a : (double)f32_unkey(block_select_key<NT>(x, n, hi, hist, bc));
    if (threadIdx.x == 0) {
        const double d = b - a;
        out[blockIdx.x] = (t >= 0.5) ? b - d * (1.0 - t) : a + d * t;
    }
}

// sums over the local chains (exact integers): out = { sum accept, sum dist, sum dist^2, n }
template <int NT>
__global__ void __launch_bounds__(NT) population_sums_kernel(const uint8_t* __restrict__ accept, const int32_t* __restrict__ dist,
                                                             int64_t n, long long* __restrict__ out) {
    __shared__ long long red[3][NT / 32];
    long long a = 0, d = 0, d2 = 0;
    for (int64_t i = threadIdx.x; i < n; i += NT) {
        if (accept) a += accept[i];
        if (dist) { const long long v = dist[i]; d += v; d2 += v * v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); d += __shfl_xor_sync(0xffffffffu, d, o); d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = d; red[2][threadIdx.x >> 5] = d2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        a = d = d2 = 0;
        for (int w = 0; w < NT / 32; ++w) { a += red[0][w]; d += red[1][w]; d2 += red[2][w]; }
        out[0] = a; out[1] = d; out[2] = d2; out[3] = n;
    }
}

// Number of DISTINCT sequences among aa[0..n): one warp per sequence inserts its index into an open-addressing table
// (int32, -1 = empty, start slot from an FNV-1a hash, linear probing).  A claimed slot is compared byte for byte with the
// inserting sequence: identical -> duplicate (done), different -> next slot.  The number of successful claims is the number of
// distinct sequences whatever the insertion order, and no hash collision can change it.
__global__ void __launch_bounds__(256) unique_count_kernel(const uint8_t* __restrict__ aa, int64_t stride, int64_t n, int L,
                                                           int32_t* __restrict__ table, uint32_t cap_mask, int32_t* __restrict__ count) {
    const int64_t b = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= n) return;
    const uint8_t* a = aa + b * stride;
    // FNV-1a over the lane's residues, lanes folded together by xor (any deterministic function of the sequence will do: the
    // hash only picks the first slot)
    unsigned long long h = 1469598103934665603ull + (unsigned long long)lane * 0x9E3779B97F4A7C15ull;
    for (int i = lane; i < L; i += 32) { h ^= a[i]; h *= 1099511628211ull; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h ^= __shfl_xor_sync(0xffffffffu, h, o);
    h *= 1099511628211ull;
    h ^= h >> 29;
    uint32_t slot = (uint32_t)(h ^ (h >> 32)) & cap_mask;
    for (;;) {
        int owner = 0;
        if (lane == 0) owner = atomicCAS(&table[slot], -1, (int32_t)b);
        owner = __shfl_sync(0xffffffffu, owner, 0);
        if (owner == -1) {                                         // first of its kind
            if (lane == 0) atomicAdd(count, 1);
            return;
        }
        const uint8_t* o = aa + (int64_t)owner * stride;
        bool diff = false;
        for (int i = lane; i < L; i += 32) diff |= (a[i] != o[i]);
        if (!__any_sync(0xffffffffu, diff)) return;               // duplicate of sequence `owner`
        slot = (slot + 1u) & cap_mask;
    }
}

// torch.topk(x, k) (largest, sorted descending; ties -> lowest index first): one CTA.
//   1. radix-select the k-th largest key;  2. in index order, take every element above it and the first (k - #above) elements
//   equal to it;  3. sort the k winners by (key desc, index asc) in shared memory (odd-even transposition, k <= 1024).
template <int NT>
__global__ void __launch_bounds__(NT) topk_kernel(const float* __restrict__ x, int64_t n, int k, int64_t index_base,
                                                  const long long* __restrict__ ids,
                                                  float* __restrict__ vals, long long* __restrict__ idx) {
    __shared__ int hist[256];
    __shared__ int64_t bc[2];
    __shared__ uint32_t skey[1024];
    __shared__ long long sidx[1024];
    __shared__ int wsum[2][NT / 32];
    __shared__ int run[2];
    const uint32_t kth = block_select_key<NT>(x, n, n - k, hist, bc);    // k-th largest = rank n-k ascending
    // how many are strictly above
    if (threadIdx.x == 0) { run[0] = 0; run[1] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // pass A: count strictly-above (needed to know how many ties to take)
    int above = 0;
    for (int64_t i = threadIdx.x; i < n; i += NT) above += (f32_key(x[i]) > kth);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) above += __shfl_xor_sync(0xffffffffu, above, o);
    if (lane == 0) wsum[0][warp] = above;
    __syncthreads();
    if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < NT / 32; ++w) s += wsum[0][w]; run[0] = s; }
    __syncthreads();
    const int n_above = run[0];
    const int ties_wanted = k - n_above;
    __syncthreads();
    if (threadIdx.x == 0) { run[0] = 0; run[1] = 0; }           // running output offsets: [0] above, [1] ties
    __syncthreads();
    // pass B: ordered compaction in index order (chunks of NT elements, block-wide exclusive scans)
    for (int64_t base = 0; base < n; base += NT) {
        const int64_t i = base + threadIdx.x;
        uint32_t key = 0u;
        int fa = 0, ft = 0;
        if (i < n) { key = f32_key(x[i]); fa = key > kth; ft = key == kth; }
        int ia = fa, it = ft;                                      // inclusive warp scans
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int va = __shfl_up_sync(0xffffffffu, ia, o), vt = __shfl_up_sync(0xffffffffu, it, o);
            if (lane >= o) { ia += va; it += vt; }
        }
        if (lane == 31) { wsum[0][warp] = ia; wsum[1][warp] = it; }
        __syncthreads();
        int offa = run[0], offt = run[1];
        for (int w = 0; w < warp; ++w) { offa += wsum[0][w]; offt += wsum[1][w]; }
        if (fa) { const int o = offa + ia - 1; skey[o] = key; sidx[o] = ids ? ids[i] : index_base + i; }
        if (ft) { const int o = offt + it - 1; if (o < ties_wanted) { skey[n_above + o] = key; sidx[n_above + o] = ids ? ids[i] : index_base + i; } }
        __syncthreads();
        if (threadIdx.x == NT - 1) { run[0] = offa + ia; run[1] = offt + it; }
        __syncthreads();
    }
    // sort (key desc, idx asc)
    for (int pass = 0; pass < k; ++pass) {
        for (int j = 2 * threadIdx.x + (pass & 1); j + 1 < k; j += 2 * NT) {
            const bool swap = skey[j] < skey[j + 1] || (skey[j] == skey[j + 1] && sidx[j] > sidx[j + 1]);
            if (swap) {
                const uint32_t tk = skey[j]; skey[j] = skey[j + 1]; skey[j + 1] = tk;
                const long long ti = sidx[j]; sidx[j] = sidx[j + 1]; sidx[j + 1] = ti;
            }
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < k; j += NT) { vals[j] = f32_unkey(skey[j]); idx[j] = sidx[j]; }
}

// rows of a residue matrix by index: out[j] = aa[idx[j] - index_base]  (sequences of the top-k chains)
__global__ void gather_rows_kernel(const uint8_t* __restrict__ aa, int64_t stride, const long long* __restrict__ idx, int k,
                                   int64_t index_base, uint8_t* __restrict__ out) {
    const int j = blockIdx.x;
    if (j >= k) return;
    const uint8_t* src = aa + (idx[j] - index_base) * stride;
    for (int64_t i = threadIdx.x; i < stride; i += blockDim.x) out[(int64_t)j * stride + i] = src[i];
}

}  // namespace ppde

using namespace ppde;

extern "C" int ppde_quantiles(const float* x, int64_t n, const double* q, int32_t nq, double* out, void* stream) {
    if (nq <= 0) return 0;
    if (!x || !q || !out || n <= 0) return (int)cudaErrorInvalidValue;
    quantile_kernel<1024><<<nq, 1024, 0, (cudaStream_t)stream>>>(x, n, q, out);
    return launch_done();
}

extern "C" int ppde_population_sums(const uint8_t* accept, const int32_t* dist, int64_t n, long long* out, void* stream) {
    if (!out || n < 0) return (int)cudaErrorInvalidValue;
    population_sums_kernel<1024><<<1, 1024, 0, (cudaStream_t)stream>>>(accept, dist, n, out);
    return launch_done();
}

extern "C" int64_t ppde_unique_count_table_entries(int64_t n) {
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}

extern "C" int ppde_unique_count(const uint8_t* aa, int64_t aa_stride, int64_t n, int32_t L, int32_t* table, int64_t table_entries,
                                 int32_t* count, void* stream) {
    if (!count) return (int)cudaErrorInvalidValue;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    if (n <= 0) return 0;
    if (!aa || !table || L <= 0 || aa_stride < L || n >= (1ll << 31) || table_entries < 2 * n ||
        (table_entries & (table_entries - 1)) || table_entries > (1ll << 32))
        return (int)cudaErrorInvalidValue;
    e = cudaMemsetAsync(table, 0xFF, (size_t)table_entries * sizeof(int32_t), st);
    if (e != cudaSuccess) return (int)e;
    unique_count_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(aa, aa_stride, n, L, table, (uint32_t)(table_entries - 1), count);
    return launch_done();
}

extern "C" int ppde_topk(const float* x, int64_t n, int32_t k, int64_t index_base, const long long* ids, float* vals,
                         long long* idx, void* stream) {
    if (k <= 0) return 0;
    if (!x || !vals || !idx || k > 1024 || k > n) return (int)cudaErrorInvalidValue;
    topk_kernel<1024><<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, k, index_base, ids, vals, idx);
    return launch_done();
}

extern "C" int ppde_gather_rows(const uint8_t* aa, int64_t aa_stride, const long long* idx, int32_t k, int64_t index_base,
                                uint8_t* out, void* stream) {
    if (k <= 0) return 0;
    if (!aa || !idx || !out) return (int)cudaErrorInvalidValue;
    gather_rows_kernel<<<k, 128, 0, (cudaStream_t)stream>>>(aa, aa_stride, idx, k, index_base, out);
    return launch_done();
}
