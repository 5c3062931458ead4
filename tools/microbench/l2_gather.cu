// Microbenchmark: per-SM throughput of gathering 960-byte rows from an L2-resident table (the access pattern of the CNN
// backward's adjoint rows), LSU path (LDG.128 into registers) vs bulk-copy path (cp.async.bulk into shared memory).
// Build:  nvcc -O3 -std=c++17 -cudart shared -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_gather l2_gather.cu (never leave the binary in the tree) ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ROWF = 240;                 // floats per row (960 B)
constexpr int NROWS = 1428;               // 3 nets x 476 rows = 1.37 MB (L2 resident)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// each warp fetches `per_warp` random rows; lane l reads 32 bytes (2 x LDG.128) of the row; U rows in flight per warp
template <int U>
__global__ void gather_ldg(const float* __restrict__ tab, int per_warp, float* __restrict__ sink) {
    const int lane = threadIdx.x & 31, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float4 acc = make_float4(0, 0, 0, 0);
    uint32_t s = hash32(gw * 977u + 1u);
    for (int i = 0; i < per_warp; i += U) {
        float4 v[U][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            s = hash32(s + u);
            const float4* p = reinterpret_cast<const float4*>(tab + (size_t)(s % NROWS) * ROWF) + 2 * lane;
            if (lane < 30) { v[u][0] = __ldg(p); v[u][1] = __ldg(p + 1); } else { v[u][0] = v[u][1] = make_float4(0, 0, 0, 0); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += v[u][0].x + v[u][1].x; acc.y += v[u][0].y + v[u][1].y; acc.z += v[u][0].z + v[u][1].z; acc.w += v[u][0].w + v[u][1].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) sink[gw] = acc.x;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each warp keeps S bulk copies (one row each) in flight into its private shared-memory slots
template <int S>
__global__ void gather_bulk(const float* __restrict__ tab, int per_warp, float* __restrict__ sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    float* slots = reinterpret_cast<float*>(smem) + (size_t)w * S * ROWF;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nw * S * ROWF * 4) + w * S;
    if (lane == 0) for (int i = 0; i < S; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t s = hash32(gw * 977u + 1u);
    auto issue = [&](int slot) {
        s = hash32(s + slot);
        if (lane == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[slot])), "r"(ROWF * 4) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(slots + slot * ROWF)),
                         "l"(tab + (size_t)(s % NROWS) * ROWF), "r"(ROWF * 4), "r"(smem_u32(&bars[slot])) : "memory");
        }
    };
    for (int i = 0; i < S && i < per_warp; ++i) issue(i);
    float acc = 0.f;
    for (int i = 0; i < per_warp; ++i) {
        const int slot = i % S;
        const uint32_t par = (i / S) & 1;
        asm volatile("{\n .reg .pred p;\nWL:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DN;\n bra WL;\nDN:\n}\n" ::"r"(smem_u32(&bars[slot])), "r"(par) : "memory");
        if (lane < 30) { const float4 a = reinterpret_cast<const float4*>(slots + slot * ROWF)[2 * lane]; acc += a.x + a.w; }
        __syncwarp();
        if (i + S < per_warp) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); issue(slot); }
    }
    if (acc == 123.456f) sink[gw] = acc;
}

int main() {
    float* tab; float* sink;
    cudaMalloc(&tab, (size_t)NROWS * ROWF * 4); cudaMemset(tab, 0, (size_t)NROWS * ROWF * 4);
    cudaMalloc(&sink, 1 << 22);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int per_warp = 4096;
    auto report = [&](const char* name, int warps, float ms) {
        const double bytes = (double)sms * warps * per_warp * ROWF * 4;
        printf("%-28s warps/SM %2d : %7.3f ms  %7.1f GB/s total  %5.1f B/clk/SM (at %.2f GHz nominal)\n", name, warps, ms, bytes / ms / 1e6,
               bytes / sms / (ms * 1e-3 * clk * 1e3), clk / 1e6);
    };
    for (int warps : {8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); gather_ldg<4><<<sms, warps * 32>>>(tab, per_warp, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) report("LDG.128, 4 rows in flight", warps, ms);
        }
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); gather_ldg<8><<<sms, warps * 32>>>(tab, per_warp, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) report("LDG.128, 8 rows in flight", warps, ms);
        }
    }
    for (int warps : {8, 16}) {
        {
            const size_t smem = (size_t)warps * 4 * ROWF * 4 + warps * 4 * 8 + 128;
            cudaFuncSetAttribute(gather_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0); gather_bulk<4><<<sms, warps * 32, smem>>>(tab, per_warp, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) report("bulk copy, 4 rows in flight", warps, ms);
            }
        }
        {
            const size_t smem = (size_t)warps * 8 * ROWF * 4 + warps * 8 * 8 + 128;
            cudaFuncSetAttribute(gather_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0); gather_bulk<8><<<sms, warps * 32, smem>>>(tab, per_warp, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) report("bulk copy, 8 rows in flight", warps, ms);
            }
        }
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
