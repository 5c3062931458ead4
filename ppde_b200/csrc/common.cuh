// Shared device helpers for the PPDE hot-path kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PPDE_Q 20                      // residue alphabet size (ppde/third_party/hsu/data_utils.py:48-72)
#define PPDE_EPS 1.1920928955078125e-07f   // torch.finfo(float32).eps used by clamp_probs
#define PPDE_MAX_S 32                  // max sub-steps per iteration (2*pas-1)

namespace ppde {

// ---------------------------------------------------------------- Philox4x32-10
// Bit-identical to ppde_b200/philox.py (Random123 known answers are tested on both sides).
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
            uint32_t n0 = hi1 ^ c1 ^ a;
            uint32_t n2 = hi0 ^ c3 ^ b;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};
// The same generator with the ten round keys precomputed on the host and passed as a kernel argument: the keys sit in the
// constant bank and feed the LOP3 of every round directly (the in-kernel key schedule was 18 integer adds per call - a quarter
// of the call - because the seed arrives in a vector register).
struct PhiloxKeys {
    uint32_t a[10], b[10];
    __host__ __device__ explicit PhiloxKeys(uint64_t seed) {
        uint32_t x = (uint32_t)seed, y = (uint32_t)(seed >> 32);
        for (int r = 0; r < 10; ++r) { a[r] = x; b[r] = y; x += 0x9E3779B9u; y += 0xBB67AE85u; }
    }
};
__device__ __forceinline__ uint4 philox_keyed(const PhiloxKeys& K, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ K.a[r];
        const uint32_t n2 = hi0 ^ c3 ^ K.b[r];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    return make_uint4(c0, c1, c2, c3);
}
enum { KIND_PROPOSAL = 0, KIND_PATHLEN = 1, KIND_ACCEPT = 2 };

// 24 random bits -> (k + 0.5) 2^-24.  For k < 2^23 the value is exact and strictly inside (0, 1); for k >= 2^23 the "+ 0.5" is
// a round-to-even tie in fp32, so the grid is the even multiples of 2^-24 there and k = 2^24 - 1 gives exactly 1.0f (probability
// 2^-24 per draw: -log(u) = 0 makes that entry's race quotient infinite-negative, i.e. it cannot win; an accept uniform of 1.0
// accepts only log_acc >= 0).  ppde_b200/philox.py has the same formula and the reference consumed these streams when the golden
// vectors were made, so the definition is part of the pinned stream; a 23-bit variant would be exact everywhere but needs the
// goldens regenerated (deliberate differences 1).
__device__ __forceinline__ float u32_to_unit(uint32_t x) {
    return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f;  // 2^-24
}

// ---------------------------------------------------------------- block reductions
// All threads must call. `red` is a __shared__ scratch of >= 33 floats (or 2x for the pair version).
template <int NT>
__device__ __forceinline__ float block_max(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    constexpr int NW = NT / 32;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) r = fmaxf(r, red[w]);
    return r;
}

template <int NT>
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    constexpr int NW = NT / 32;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) r += red[w];     // fixed order: deterministic
    return r;
}

template <int NT>
__device__ __forceinline__ int block_sum_int(int v, int* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    constexpr int NW = NT / 32;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) r += red[w];
    return r;
}

// argmax with lowest-index tie break (torch.argmax / torch.max on CPU return the first maximum).
template <int NT>
__device__ __forceinline__ void block_argmax(float& v, int& idx, float* redv, int* redi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, v, o);
        int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    constexpr int NW = NT / 32;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = v; redi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    v = redv[0]; idx = redi[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) {
        float ov = redv[w]; int oi = redi[w];
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

__device__ __forceinline__ float clamp_prob(float p) {          // torch.distributions.utils.clamp_probs
    return fminf(fmaxf(p, PPDE_EPS), 1.0f - PPDE_EPS);
}

}  // namespace ppde
