// CNN-ensemble forward on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
// Computes, for every chain b, net k and hidden channel j (reference: OnehotCNN.forward,
// ppde/nets.py:363-376):   m[j] = max_p relu(b1[j] + sum_c W1[j,c] * r1[p,c]),  p*_j = lowest arg-max
// with r1[p,c] = relu(b0[c] + sum_{t<5} W0[c, aa[p+t], t])  (one-hot conv = 5 table gathers).
//
// GEMM view (per net):  D[j, (b,p)] = W1[j, :] . r1[(b,p), :]     M = 2C channels, N = positions, K = C.
//   * A = W1 tile (128 channels x K) lives in TENSOR MEMORY for the whole kernel (fp16 hi + lo halves,
//     written once with tcgen05.st) -> the MMA reads no shared memory for A.
//   * B = r1 tile (N_tile positions x 64-wide K chunk) is PRODUCED on the fly by 8 producer warps from a
//     shared-memory copy of the conv table (b0 folded in), split into fp16 hi/lo, and stored K-major with the
//     128-byte swizzle the UMMA descriptor expects; a 3-slot ring of chunks pipelines producers and MMA.
//   * fp32 parity: x*y ~ xh*yh + xh*yl + xl*yh with fp32 accumulation in TMEM (3 fp16 MMAs per K step).
//     fp16 carries 11 significand bits, so hi+lo holds 22 bits and the dropped xl*yl term is ~2^-22 relative;
//     both operands are pre-scaled by per-net powers of two (w1_scale, r1_scale; undone exactly in the epilogue)
//     so that the lo halves stay in fp16's normal range.
//   * D (128 channels x N_tile positions, fp32) is double-buffered in TMEM; 4 epilogue warps read it with
//     tcgen05.ld (thread = channel), add bias, relu, and keep a running (max, first arg-max) in registers
//     across the tiles of a chain, then write the 64-bit winner key.  No atomics, no memset.
// Warp roles: warps 0-3 epilogue (TMEM lane quarters), warp 4 MMA issuer (one elected lane), warps 5-20 producers.
// Persistent grid: CTA -> (net, channel tile) x contiguous block of chains.
#include "common.cuh"
#include "../../include/ppde_b200.h"
#include "launch.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace ppde {
namespace tc {

constexpr int NT_EPI = 128;
constexpr int NT_PROD = 512;                       // 16 producer warps (4 per scheduler): the producers are issue/latency bound
constexpr int NTHREADS = NT_EPI + 32 + NT_PROD;   // 672
constexpr int WARP_MMA = 4;
constexpr int NSLOT = 3;
constexpr int KCH = 64;                            // K elements per chunk = one 128-byte swizzle row of fp16
constexpr int MAT_BYTES = 128 * KCH * 2;           // one [128 x 64] fp16 operand matrix (16 KB)
constexpr int SLOT_BYTES = 2 * MAT_BYTES;          // hi + lo
constexpr int TMEM_COLS = 512;
constexpr int D_COL0 = 256;                        // accumulators: columns [256,384) and [384,512)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "WAIT_LOOP:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE;\n"
        " bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]   (A from tensor memory, K-major; f16 x f16 -> fp32)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// K-major, 128-byte-swizzled operand matrix: rows of 64 halves (128 B), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    // c=f32 (bit4), a=f16 (bits 7..9 = 0), b=f16 (bits 10..12 = 0), both K-major, N>>3 at 17, M>>4 at 24
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_h2(float lo_k, float hi_k) {       // low 16 bits = even k
    __half2 v = __floats2half2_rn(lo_k, hi_k);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float h_round(float x) { return __half2float(__float2half_rn(x)); }
// packed fp32 add (FADD2 on sm_100): halves the issue slots of the 5-tap table sums
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return r;
}
// x (>= 0, pre-scaled) -> fp16 hi (top 11 significand bits, by truncation: exact in fp16) and the exact fp32 residual
__device__ __forceinline__ float h_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

struct Params {
    ppde_cnn_t m;
    const uint8_t* aa;
    int aa_stride;
    int n;
    unsigned long long* mkey;
    uint8_t* r1mask;      // optional [n, n_nets, P, 32]: bit c of a position's 32 bytes = (r1[p,c] > 0), for the backward
    int n_tile;           // positions per tile (multiple of 16, <= 128)
    int tiles_per_chain;
    int ctas_per_combo;
    int MT;               // channel tiles of 128
    int nch;              // K chunks of 64
    int kpad;             // K rounded up to 16
};

__global__ void __launch_bounds__(NTHREADS, 1) cnn_forward_tc_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int C = prm.m.C, P = prm.m.P, J2 = 2 * C;
    const int KS = prm.nch * KCH;                                     // padded table row length
    // NSLOT x (hi 16 KB | lo 16 KB); the 128-byte swizzle is a function of address bits, so align to 1024 B
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* sT0 = reinterpret_cast<float*>(ring + NSLOT * SLOT_BYTES);   // [100][KS], b0 folded into tap 0
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT0 + 100 * KS);
    uint64_t* full = bars;              // [NSLOT] producers -> MMA
    uint64_t* empty = bars + NSLOT;     // [NSLOT] MMA -> producers
    uint64_t* dfull = empty + NSLOT;    // [2]     MMA -> epilogue
    uint64_t* dempty = dfull + 2;       // [2]     epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

    const int combo = blockIdx.x / prm.ctas_per_combo;
    const int within = blockIdx.x - combo * prm.ctas_per_combo;
    if (combo >= prm.m.n_nets * prm.MT) return;
    const int k = combo / prm.MT, mt = combo - k * prm.MT;
    const ppde_cnn_net_t net = prm.m.net[k];
    const int b_lo = (int)((int64_t)prm.n * within / prm.ctas_per_combo);
    const int b_hi = (int)((int64_t)prm.n * (within + 1) / prm.ctas_per_combo);
    const int ntiles = (b_hi - b_lo) * prm.tiles_per_chain;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup ---------------------------------------------------------------------
    for (int e = threadIdx.x; e < 100 * KS; e += NTHREADS) {
        const int row = e / KS, c = e - row * KS;
        float v = 0.f;
        if (c < C) {
            v = net.T0[(size_t)row * C + c];
            if (row < PPDE_Q) v += net.b0[c];                         // tap 0 rows carry the bias
        }
        sT0[e] = v * net.r1_scale;                                    // power of two: exact; r1 comes out pre-scaled
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT; ++s) { mbar_init(&full[s], NT_PROD / 32); mbar_init(&empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&dfull[d], 1); mbar_init(&dempty[d], NT_EPI); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // A = W1 rows [mt*128, +128) -> TMEM: lane = channel, 32-bit column = two consecutive k (fp16 hi at
        // columns [0, kpad/2), residual lo at [kpad/2, kpad)).
        const int j = mt * 128 + warp * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int ks = 0; ks < prm.kpad / 16; ++ks) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c0 = ks * 16 + 2 * q;
                const float w0 = (j < J2 && c0 < C) ? net.W1[(size_t)j * C + c0] * net.w1_scale : 0.f;
                const float w1 = (j < J2 && c0 + 1 < C) ? net.W1[(size_t)j * C + c0 + 1] * net.w1_scale : 0.f;
                const float h0 = h_round(w0), h1 = h_round(w1);
                hi[q] = pack_h2(h0, h1);
                lo[q] = pack_h2(w0 - h0, w1 - h1);
            }
            tmem_st8(lane_addr + ks * 8, hi);
            tmem_st8(lane_addr + prm.kpad / 2 + ks * 8, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- role dispatch --------------------------------------------------------------------------
    if (warp < 4) {
        // ===== EPILOGUE: thread = channel j =====
        const int j = mt * 128 + warp * 32 + lane;
        const float bias = (j < J2) ? net.b1[j] : 0.f;
        const float unscale = 1.f / (net.w1_scale * net.r1_scale);     // exact: both scales are powers of two
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL0;
        float best = -1.f;
        int bp = 0;
        for (int it = 0; it < ntiles; ++it) {
            const int buf = it & 1;
            const int b = b_lo + it / prm.tiles_per_chain;
            const int tn = it - (it / prm.tiles_per_chain) * prm.tiles_per_chain;
            const int p0 = tn * prm.n_tile;
            const int valid = min(prm.n_tile, P - p0);
            if (tn == 0) { best = -1.f; bp = 0; }
            mbar_wait(&dfull[buf], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            for (int cg = 0; cg * 32 < valid; ++cg) {
                uint32_t r[32];
                tmem_ld32(lane_addr + buf * 128 + cg * 32, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float v = fmaf(__uint_as_float(r[i]), unscale, bias);
                    v = v > 0.f ? v : 0.f;
                    if (cg * 32 + i < valid && v > best) { best = v; bp = p0 + cg * 32 + i; }
                }
            }
            tc_fence_before();
            mbar_arrive(&dempty[buf]);
            if (tn == prm.tiles_per_chain - 1 && j < J2) {
                const unsigned long long key =
                    ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)bp);
                prm.mkey[((size_t)b * prm.m.n_nets + k) * J2 + j] = key;
            }
        }
    } else if (warp == WARP_MMA) {
        // ===== MMA ISSUER (one lane) =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(128, prm.n_tile);
            const uint32_t ring_addr = smem_u32(ring);
            int slot = 0;
            uint32_t sphase = 0;
            const int last_ksteps = (prm.kpad - (prm.nch - 1) * KCH) / 16;
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it & 1;
                if (it >= 2) mbar_wait(&dempty[buf], (uint32_t)(((it >> 1) + 1) & 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + D_COL0 + buf * 128;
                for (int kc = 0; kc < prm.nch; ++kc) {
                    mbar_wait(&full[slot], sphase);
                    tc_fence_after();
                    const uint64_t dhi = make_b_desc(ring_addr + slot * SLOT_BYTES);
                    const uint64_t dlo = make_b_desc(ring_addr + slot * SLOT_BYTES + MAT_BYTES);
                    const int ksteps = (kc == prm.nch - 1) ? last_ksteps : KCH / 16;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t a_hi = tmem_base + kc * (KCH / 2) + ks * 8;
                        const uint32_t a_lo = a_hi + prm.kpad / 2;
                        const uint64_t koff = (uint64_t)(ks * 2);                 // +32 bytes per K step (>>4)
                        mma_ts(d_tmem, a_hi, dhi + koff, idesc, (kc | ks) ? 1u : 0u);
                        mma_ts(d_tmem, a_hi, dlo + koff, idesc, 1u);
                        mma_ts(d_tmem, a_lo, dhi + koff, idesc, 1u);
                    }
                    tc_commit(&empty[slot]);                                      // frees the ring slot when the MMAs retire
                    if (++slot == NSLOT) { slot = 0; sphase ^= 1; }
                }
                tc_commit(&dfull[buf]);                                           // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== PRODUCERS: r1 chunk -> fp16 hi/lo, K-major SW128 =====
        // lane = g + 8q: g = channel group (channels cb+4g..+3 and cb+32+4g..+3: conflict-free 128-byte LDS phases),
        // q = one of the warp's 4 rows; rows {x, x+4, x+8, x+12} per warp keep the 8-byte swizzled stores conflict-free.
        const int pw = warp - 5;                         // 0..15
        const int g = lane & 7, q = lane >> 3;
        const int rsub = 16 * (pw >> 2) + (pw & 3) + 4 * q;   // row inside a 64-row pass
        int slot = 0;
        uint32_t phase = 0;
        for (int it = 0; it < ntiles; ++it) {
            const int b = b_lo + it / prm.tiles_per_chain;
            const int tn = it - (it / prm.tiles_per_chain) * prm.tiles_per_chain;
            const int p0 = tn * prm.n_tile;
            const int valid = min(prm.n_tile, P - p0);
            const uint8_t* a = prm.aa + (size_t)b * prm.aa_stride + p0;
            const float* trow[2][5];                       // table rows (t*20 + aa[p+t]) of my 2 rows, at my channel group
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = 64 * i + rsub;
#pragma unroll
                for (int t = 0; t < 5; ++t)
                    trow[i][t] = sT0 + ((r < valid) ? (t * PPDE_Q + a[r + t]) * KS : 0) + 4 * g;
            }
            // the MT CTAs that share a chain block produce identical r1 tiles: they take turns writing the relu mask
            const bool emit_mask = (prm.r1mask != nullptr) && (it % prm.MT == mt);
            uint32_t mbits[2] = {0u, 0u};                  // nibble 2*kc + h = (r1 > 0) of channels kc*64 + 32h + 4g ..
            for (int kc = 0; kc < prm.nch; ++kc) {
                mbar_wait(&empty[slot], phase ^ 1);
                unsigned char* mat_hi = ring + slot * SLOT_BYTES;
                unsigned char* mat_lo = mat_hi + MAT_BYTES;
                const int cb = kc * KCH;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = 64 * i + rsub;
                    if (r < valid) {
                        float4 z0 = *reinterpret_cast<const float4*>(trow[i][0] + cb);
                        float4 z1 = *reinterpret_cast<const float4*>(trow[i][0] + cb + 32);
                        float2 a0 = make_float2(z0.x, z0.y), a1 = make_float2(z0.z, z0.w);
                        float2 a2 = make_float2(z1.x, z1.y), a3 = make_float2(z1.z, z1.w);
#pragma unroll
                        for (int t = 1; t < 5; ++t) {
                            const float4 u0 = *reinterpret_cast<const float4*>(trow[i][t] + cb);
                            const float4 u1 = *reinterpret_cast<const float4*>(trow[i][t] + cb + 32);
                            a0 = add2(a0, make_float2(u0.x, u0.y)); a1 = add2(a1, make_float2(u0.z, u0.w));
                            a2 = add2(a2, make_float2(u1.x, u1.y)); a3 = add2(a3, make_float2(u1.z, u1.w));
                        }
                        if (emit_mask) {                   // warp-uniform: only every MT-th tile writes mask bits
                            const float v[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
                            uint32_t nib = 0u;
#pragma unroll
                            for (int e = 0; e < 8; ++e) nib |= (__float_as_int(v[e]) > 0 ? 1u : 0u) << e;
                            mbits[i] |= nib << (8 * kc);
                            }
                        // relu, fp16 hi by truncation, exact residual (packed), pack

                        float2 x[4] = {a0, a1, a2, a3};
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            x[e].x = fmaxf(x[e].x, 0.f); x[e].y = fmaxf(x[e].y, 0.f);
                            const float2 h = make_float2(h_trunc(x[e].x), h_trunc(x[e].y));
                            const float2 l = add2(x[e], make_float2(-h.x, -h.y));
                            hi[e] = pack_h2(h.x, h.y);
                            lo[e] = pack_h2(l.x, l.y);
                        }

                        // element (row r, k) at (r/8)*1024 + (r%8)*128 + ((k/8) ^ (r%8))*16 + (k%8)*2
                        const int rbase = (r >> 3) * 1024 + (r & 7) * 128;
                        const int o0 = rbase + (((g >> 1) ^ (r & 7)) << 4) + ((g & 1) << 3);          // k = 4g
                        const int o1 = rbase + ((((g >> 1) + 4) ^ (r & 7)) << 4) + ((g & 1) << 3);    // k = 32 + 4g
                        *reinterpret_cast<uint2*>(mat_hi + o0) = make_uint2(hi[0], hi[1]);
                        *reinterpret_cast<uint2*>(mat_hi + o1) = make_uint2(hi[2], hi[3]);
                        *reinterpret_cast<uint2*>(mat_lo + o0) = make_uint2(lo[0], lo[1]);
                        *reinterpret_cast<uint2*>(mat_lo + o1) = make_uint2(lo[2], lo[3]);
                    }
                }
                fence_proxy_async();                      // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[slot]);
                if (++slot == NSLOT) { slot = 0; phase ^= 1; }
            }
            if (emit_mask) {
                // word w of a row's 256 mask bits = channels 32w..32w+31 = nibble w of the row's 8 lanes:
                // 8x8 nibble transpose across those lanes (3 butterfly stages), then one coalesced 32-byte store per row
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = 64 * i + rsub;
                    uint32_t word = mbits[i];
                    uint32_t o = __shfl_xor_sync(0xffffffffu, word, 4);
                    word = (g & 4) ? ((word & 0xFFFF0000u) | ((o >> 16) & 0x0000FFFFu)) : ((word & 0x0000FFFFu) | ((o << 16) & 0xFFFF0000u));
                    o = __shfl_xor_sync(0xffffffffu, word, 2);
                    word = (g & 2) ? ((word & 0xFF00FF00u) | ((o >> 8) & 0x00FF00FFu)) : ((word & 0x00FF00FFu) | ((o << 8) & 0xFF00FF00u));
                    o = __shfl_xor_sync(0xffffffffu, word, 1);
                    word = (g & 1) ? ((word & 0xF0F0F0F0u) | ((o >> 4) & 0x0F0F0F0Fu)) : ((word & 0x0F0F0F0Fu) | ((o << 4) & 0xF0F0F0F0u));
                    if (r < valid)
                        reinterpret_cast<uint32_t*>(prm.r1mask + (((size_t)b * prm.m.n_nets + k) * P + p0 + r) * 32)[g] = word;
                }
            }
        }
    }

    // ---- teardown -----------------------------------------------------------------------------
    tc_fence_before();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}


// =====================================================================================================
// 2-CTA variant of the forward kernel (cta_group::2): a cluster of two CTAs (one TPC) computes M = 256 channels
// (two channel tiles) per MMA.  Each CTA keeps ITS 128-channel W1 tile in its own tensor memory and PRODUCES only
// half of the positions of every tile (the B operand of a cta_group::2 MMA is split along N across the pair), so the
// r1 production cost per MMA flop — the bottleneck of the 1-CTA kernel (4x redundant across channel tiles) — halves,
// and so does the shared-memory traffic of the B operand per CTA.  The leader CTA (rank 0) issues the MMAs;
// `full` / `dempty` barriers live in the leader and collect arrivals from both CTAs (mapa + cluster-scope arrive),
// `empty` / `dfull` are signalled in both CTAs by multicast tcgen05.commit.
constexpr int NSLOT2 = 6;
constexpr int MAT2_BYTES = 64 * KCH * 2;          // one [64 x 64] fp16 operand half-matrix (8 KB)
constexpr int SLOT2_BYTES = 2 * MAT2_BYTES;       // hi + lo

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release, cluster scope) on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "WAIT_LOOP_C:\n"
        " mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_C;\n"
        " bra WAIT_LOOP_C;\n"
        "DONE_C:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {        // both CTAs' barrier at this offset
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
cnn_forward_tc2_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int C = prm.m.C, P = prm.m.P, J2 = 2 * C;
    const int KS = prm.nch * KCH;
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* sT0 = reinterpret_cast<float*>(ring + NSLOT2 * SLOT2_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT0 + 100 * KS);
    uint64_t* full = bars;                 // [NSLOT2] (leader's copy is the live one)
    uint64_t* empty = bars + NSLOT2;       // [NSLOT2] local
    uint64_t* dfull = empty + NSLOT2;      // [2] local
    uint64_t* dempty = dfull + 2;          // [2] (leader's copy is the live one)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int MP = prm.MT;                                   // here: number of channel-tile PAIRS
    const int combo = pair / prm.ctas_per_combo;             // ctas_per_combo = pairs per (net, channel-tile pair)
    const int within = pair - combo * prm.ctas_per_combo;
    const bool idle = combo >= prm.m.n_nets * MP;            // whole cluster idle (uniform over the pair)
    const int k = idle ? 0 : combo / MP, mp = idle ? 0 : combo - k * MP;
    const int mt = mp * 2 + (int)rank;
    const ppde_cnn_net_t net = prm.m.net[k];
    const int b_lo = idle ? 0 : (int)((int64_t)prm.n * within / prm.ctas_per_combo);
    const int b_hi = idle ? 0 : (int)((int64_t)prm.n * (within + 1) / prm.ctas_per_combo);
    const int ntiles = (b_hi - b_lo) * prm.tiles_per_chain;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_half = prm.n_tile >> 1;

    for (int e = threadIdx.x; e < 100 * KS; e += NTHREADS) {
        const int row = e / KS, c = e - row * KS;
        float v = 0.f;
        if (c < C) {
            v = net.T0[(size_t)row * C + c];
            if (row < PPDE_Q) v += net.b0[c];
        }
        sT0[e] = v * net.r1_scale;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT2; ++s) { mbar_init(&full[s], 2 * (NT_PROD / 32)); mbar_init(&empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&dfull[d], 1); mbar_init(&dempty[d], 2 * NT_EPI); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                    // barrier inits and TMEM allocation visible to the peer
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        const int j = mt * 128 + warp * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int ks = 0; ks < prm.kpad / 16; ++ks) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c0 = ks * 16 + 2 * q;
                const float w0 = (j < J2 && c0 < C) ? net.W1[(size_t)j * C + c0] * net.w1_scale : 0.f;
                const float w1 = (j < J2 && c0 + 1 < C) ? net.W1[(size_t)j * C + c0 + 1] * net.w1_scale : 0.f;
                const float h0 = h_round(w0), h1 = h_round(w1);
                hi[q] = pack_h2(h0, h1);
                lo[q] = pack_h2(w0 - h0, w1 - h1);
            }
            tmem_st8(lane_addr + ks * 8, hi);
            tmem_st8(lane_addr + prm.kpad / 2 + ks * 8, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                    // both CTAs' A tiles are in place before the leader issues any MMA
    tc_fence_after();

    if (warp < 4) {
        // ===== EPILOGUE (both CTAs): thread = channel j of this CTA's tile =====
        const int j = mt * 128 + warp * 32 + lane;
        const float bias = (j < J2) ? net.b1[j] : 0.f;
        const float unscale = 1.f / (net.w1_scale * net.r1_scale);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL0;
        float best = -1.f;
        int bp = 0;
        for (int it = 0; it < ntiles; ++it) {
            const int buf = it & 1;
            const int b = b_lo + it / prm.tiles_per_chain;
            const int tn = it - (it / prm.tiles_per_chain) * prm.tiles_per_chain;
            const int p0 = tn * prm.n_tile;
            const int valid = min(prm.n_tile, P - p0);
            if (tn == 0) { best = -1.f; bp = 0; }
            mbar_wait(&dfull[buf], (uint32_t)((it >> 1) & 1));
            tc_fence_after();
            for (int cg = 0; cg * 32 < valid; ++cg) {
                uint32_t r[32];
                tmem_ld32(lane_addr + buf * 128 + cg * 32, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float v = fmaf(__uint_as_float(r[i]), unscale, bias);
                    v = v > 0.f ? v : 0.f;
                    if (cg * 32 + i < valid && v > best) { best = v; bp = p0 + cg * 32 + i; }
                }
            }
            tc_fence_before();
            mbar_arrive_cluster(&dempty[buf], 0);
            if (tn == prm.tiles_per_chain - 1 && j < J2) {
                const unsigned long long key =
                    ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)bp);
                prm.mkey[((size_t)b * prm.m.n_nets + k) * J2 + j] = key;
            }
        }
    } else if (warp == WARP_MMA) {
        // ===== MMA ISSUER: one lane of the LEADER CTA =====
        if (rank == 0 && lane == 0) {
            const uint32_t idesc = make_idesc(256, prm.n_tile);
            const uint32_t ring_addr = smem_u32(ring);
            int slot = 0;
            uint32_t sphase = 0;
            const int last_ksteps = (prm.kpad - (prm.nch - 1) * KCH) / 16;
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it & 1;
                if (it >= 2) mbar_wait_cluster(&dempty[buf], (uint32_t)(((it >> 1) + 1) & 1));
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + D_COL0 + buf * 128;
                for (int kc = 0; kc < prm.nch; ++kc) {
                    mbar_wait_cluster(&full[slot], sphase);
                    tc_fence_after();
                    const uint64_t dhi = make_b_desc(ring_addr + slot * SLOT2_BYTES);
                    const uint64_t dlo = make_b_desc(ring_addr + slot * SLOT2_BYTES + MAT2_BYTES);
                    const int ksteps = (kc == prm.nch - 1) ? last_ksteps : KCH / 16;
                    for (int ks = 0; ks < ksteps; ++ks) {
                        const uint32_t a_hi = tmem_base + kc * (KCH / 2) + ks * 8;
                        const uint32_t a_lo = a_hi + prm.kpad / 2;
                        const uint64_t koff = (uint64_t)(ks * 2);
                        mma_ts2(d_tmem, a_hi, dhi + koff, idesc, (kc | ks) ? 1u : 0u);
                        mma_ts2(d_tmem, a_hi, dlo + koff, idesc, 1u);
                        mma_ts2(d_tmem, a_lo, dhi + koff, idesc, 1u);
                    }
                    tc_commit2(&empty[slot]);
                    if (++slot == NSLOT2) { slot = 0; sphase ^= 1; }
                }
                tc_commit2(&dfull[buf]);
            }
        }
    } else {
        // ===== PRODUCERS (both CTAs): rows [rank*n_half, (rank+1)*n_half) of every tile, one row per thread =====
        const int pw = warp - 5;
        const int g = lane & 7, q = lane >> 3;
        const int r = 16 * (pw >> 2) + (pw & 3) + 4 * q;      // local row 0..63
        int slot = 0;
        uint32_t phase = 0;
        for (int it = 0; it < ntiles; ++it) {
            const int b = b_lo + it / prm.tiles_per_chain;
            const int tn = it - (it / prm.tiles_per_chain) * prm.tiles_per_chain;
            const int p0 = tn * prm.n_tile + (int)rank * n_half;           // first position of MY half
            const int valid = min(n_half, P - p0);                         // may be <= 0
            const bool ok = r < valid;
            const uint8_t* a = prm.aa + (size_t)b * prm.aa_stride + p0;
            const float* trow[5];
#pragma unroll
            for (int t = 0; t < 5; ++t) trow[t] = sT0 + (ok ? (t * PPDE_Q + a[r + t]) * KS : 0) + 4 * g;
            const bool emit_mask = (prm.r1mask != nullptr) && (it % MP == mp);
            uint32_t mbits = 0u;
            for (int kc = 0; kc < prm.nch; ++kc) {
                mbar_wait(&empty[slot], phase ^ 1);
                unsigned char* mat_hi = ring + slot * SLOT2_BYTES;
                unsigned char* mat_lo = mat_hi + MAT2_BYTES;
                const int cb = kc * KCH;
                if (ok) {
                    float4 z0 = *reinterpret_cast<const float4*>(trow[0] + cb);
                    float4 z1 = *reinterpret_cast<const float4*>(trow[0] + cb + 32);
                    float2 a0 = make_float2(z0.x, z0.y), a1 = make_float2(z0.z, z0.w);
                    float2 a2 = make_float2(z1.x, z1.y), a3 = make_float2(z1.z, z1.w);
#pragma unroll
                    for (int t = 1; t < 5; ++t) {
                        const float4 u0 = *reinterpret_cast<const float4*>(trow[t] + cb);
                        const float4 u1 = *reinterpret_cast<const float4*>(trow[t] + cb + 32);
                        a0 = add2(a0, make_float2(u0.x, u0.y)); a1 = add2(a1, make_float2(u0.z, u0.w));
                        a2 = add2(a2, make_float2(u1.x, u1.y)); a3 = add2(a3, make_float2(u1.z, u1.w));
                    }
                    if (emit_mask) {
                        const float v[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
                        uint32_t nib = 0u;
#pragma unroll
                        for (int e = 0; e < 8; ++e) nib |= (__float_as_int(v[e]) > 0 ? 1u : 0u) << e;
                        mbits |= nib << (8 * kc);
                    }
                    float2 x[4] = {a0, a1, a2, a3};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        x[e].x = fmaxf(x[e].x, 0.f); x[e].y = fmaxf(x[e].y, 0.f);
                        const float2 h = make_float2(h_trunc(x[e].x), h_trunc(x[e].y));
                        const float2 l = add2(x[e], make_float2(-h.x, -h.y));
                        hi[e] = pack_h2(h.x, h.y);
                        lo[e] = pack_h2(l.x, l.y);
                    }
                    const int rbase = (r >> 3) * 1024 + (r & 7) * 128;
                    const int o0 = rbase + (((g >> 1) ^ (r & 7)) << 4) + ((g & 1) << 3);
                    const int o1 = rbase + ((((g >> 1) + 4) ^ (r & 7)) << 4) + ((g & 1) << 3);
                    *reinterpret_cast<uint2*>(mat_hi + o0) = make_uint2(hi[0], hi[1]);
                    *reinterpret_cast<uint2*>(mat_hi + o1) = make_uint2(hi[2], hi[3]);
                    *reinterpret_cast<uint2*>(mat_lo + o0) = make_uint2(lo[0], lo[1]);
                    *reinterpret_cast<uint2*>(mat_lo + o1) = make_uint2(lo[2], lo[3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&full[slot], 0);
                if (++slot == NSLOT2) { slot = 0; phase ^= 1; }
            }
            if (emit_mask) {
                uint32_t word = mbits;
                uint32_t o = __shfl_xor_sync(0xffffffffu, word, 4);
                word = (g & 4) ? ((word & 0xFFFF0000u) | ((o >> 16) & 0x0000FFFFu)) : ((word & 0x0000FFFFu) | ((o << 16) & 0xFFFF0000u));
                o = __shfl_xor_sync(0xffffffffu, word, 2);
                word = (g & 2) ? ((word & 0xFF00FF00u) | ((o >> 8) & 0x00FF00FFu)) : ((word & 0x00FF00FFu) | ((o << 8) & 0xFF00FF00u));
                o = __shfl_xor_sync(0xffffffffu, word, 1);
                word = (g & 1) ? ((word & 0xF0F0F0F0u) | ((o >> 4) & 0x0F0F0F0Fu)) : ((word & 0x0F0F0F0Fu) | ((o << 4) & 0xF0F0F0F0u));
                if (ok) reinterpret_cast<uint32_t*>(prm.r1mask + (((size_t)b * prm.m.n_nets + k) * P + p0 + r) * 32)[g] = word;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                    // the peer may still be reading my smem / signalling my barriers
    if (warp == WARP_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// =====================================================================================================
// Backward of the CNN ensemble on the tensor cores (same skeleton as the forward kernel).
//
//   dfit_k/dx[i,a] = sum_t sum_c A[i-t,c] W0[c,a,t],   A[p,c] = 1[r1[p,c]>0] * sum_{j: p*_j=p, m_j>0} d_j W1[j,c]
//
// GEMM view (per net):  Y[(t,a), p] = sum_c W0[c,a,t] * A[p,c]      M = 100 (t,a) rows (padded to 128), N = positions, K = C.
//   * A-operand  = W0^T tile (100 x K), fp16 hi/lo, resident in TENSOR MEMORY per net.
//   * B-operand  = adjoint rows, PRODUCED on the fly: the producers bucket the <= 2C arg-max winners of the chain
//     by position (counting sort in shared memory), recompute the relu mask of r1 from the conv table, gather-sum
//     the winners' W1 rows from L2 (coalesced 256-byte row segments) and write fp16 hi/lo K-major SW128 chunks.
//   * Epilogue   = tcgen05.ld (thread = (t,a)), deterministic col2im through a small smem tile into the chain's
//     [20L] gradient accumulator, flushed into the pool row:  G = Gp(window) + lamda/n_nets * sum_k dfit_k/dx.
// CTA -> (net, contiguous block of chains); per-net partial gradients go to a scratch buffer and a streaming
// kernel forms  G = Gp(window) + lamda/n_nets * (Gc_0 + Gc_1 + Gc_2)  in a fixed order (deterministic).
constexpr int BW_NSLOT = 3;
constexpr int BW_NT_PROD = 512;        // 16 producer warps: the W1 gathers are L2-latency bound, more warps = more loads in flight
constexpr int BW_NTHREADS = NT_EPI + 32 + BW_NT_PROD;   // 672
constexpr int YS = 33;                 // sY row stride (floats)

struct BwdParams {
    ppde_cnn_t m;
    ppde_potts_t pm;
    const uint8_t* aa;
    int aa_stride;
    int n;
    const unsigned long long* mkey;
    const uint8_t* r1mask;              // [n, n_nets, P, 32] relu mask bits written by the forward kernel
    float* Gc;                          // [n_nets][n][20L] per-net partial gradients (combined by cnn_grad_combine_kernel)
    int ctas_per_net;
    int n_tile, tiles_per_chain, nch, kpad;
};

__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(BW_NTHREADS, 1) cnn_backward_tc_kernel(const __grid_constant__ BwdParams prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int C = prm.m.C, P = prm.m.P, L = prm.m.L, J2 = 2 * C, NE = L * PPDE_Q;
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* sGc = reinterpret_cast<float*>(ring + BW_NSLOT * SLOT_BYTES);    // [NE] chain accumulator
    float* sY = sGc + NE;                                              // [100][YS]
    float* sDj = sY + 100 * YS;                                        // [J2] d_j of active winners
    int* sPst = reinterpret_cast<int*>(sDj + J2);                      // [J2] p*_j or -1
    int* sStart = sPst + J2;                                           // [P+1]
    int* sFill = sStart + (P + 1);                                     // [P]
    int* sList = sFill + P;                                            // [J2]
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sList + J2) + 7) & ~(uintptr_t)7);
    uint64_t* full = bars;
    uint64_t* empty = bars + BW_NSLOT;
    uint64_t* dfull = empty + BW_NSLOT;
    uint64_t* dempty = dfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

    const int k = blockIdx.x / prm.ctas_per_net;                       // this CTA's net
    const int within = blockIdx.x - k * prm.ctas_per_net;
    if (k >= prm.m.n_nets) return;
    const int b_lo = (int)((int64_t)prm.n * within / prm.ctas_per_net);
    const int b_hi = (int)((int64_t)prm.n * (within + 1) / prm.ctas_per_net);
    const int ntiles = (b_hi - b_lo) * prm.tiles_per_chain;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < BW_NSLOT; ++s) { mbar_init(&full[s], BW_NT_PROD / 32); mbar_init(&empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&dfull[d], 1); mbar_init(&dempty[d], NT_EPI); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int e = threadIdx.x; e < NE; e += BW_NTHREADS) sGc[e] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // pipeline state persists across the nets
    int p_slot = 0; uint32_t p_phase = 0;      // producers
    int m_slot = 0; uint32_t m_phase = 0;      // MMA issuer
    const int git = 0;

    {
        const ppde_cnn_net_t net = prm.m.net[k];
        // ---- per-net setup: W0^T (scaled, fp16 hi/lo) -> TMEM ----
        if (warp < 4) {
            const int nrow = warp * 32 + lane;                           // (t,a) = (nrow / 20, nrow % 20)
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
            for (int ks = 0; ks < prm.kpad / 16; ++ks) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int c0 = ks * 16 + 2 * q;
                    const float w0 = (nrow < 100 && c0 < C) ? net.W0r[(size_t)c0 * 100 + nrow] * net.w0_scale : 0.f;
                    const float w1 = (nrow < 100 && c0 + 1 < C) ? net.W0r[(size_t)(c0 + 1) * 100 + nrow] * net.w0_scale : 0.f;
                    const float h0 = h_round(w0), h1 = h_round(w1);
                    hi[q] = pack_h2(h0, h1);
                    lo[q] = pack_h2(w0 - h0, w1 - h1);
                }
                tmem_st8(lane_addr + ks * 8, hi);
                tmem_st8(lane_addr + prm.kpad / 2 + ks * 8, lo);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();

        if (warp < 4) {
            // ===== EPILOGUE: thread = output row (t,a) =====
            const int nrow = warp * 32 + lane;
            const int tid = threadIdx.x;                                  // 0..127
            const float unscale = 1.f / (net.w0_scale * net.adj_scale);
            const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL0;
            for (int it = 0; it < ntiles; ++it) {
                const int gt = git + it;
                const int buf = gt & 1;
                const int b = b_lo + it / prm.tiles_per_chain;
                const int tn = it - (it / prm.tiles_per_chain) * prm.tiles_per_chain;
                const int p0 = tn * prm.n_tile;
                const int valid = min(prm.n_tile, P - p0);
                mbar_wait(&dfull[buf], (uint32_t)((gt >> 1) & 1));
                tc_fence_after();
                for (int cg = 0; cg * 32 < valid; ++cg) {
                    uint32_t r[32];
                    tmem_ld32(lane_addr + buf * 128 + cg * 32, r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if ((cg + 1) * 32 >= valid) {                         // last read of this accumulator
                        tc_fence_before();
                        mbar_arrive(&dempty[buf]);
                    }
                    if (nrow < 100) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sY[nrow * YS + i] = __uint_as_float(r[i]) * unscale;
                    }
                    named_bar(2, NT_EPI);
                    // col2im, fixed summation order: output (i,a) = sum_t Y[(t,a), i - t]
                    const int nv = min(32, valid - cg * 32);
                    for (int o = tid; o < 36 * PPDE_Q; o += NT_EPI) {
                        const int di = o / PPDE_Q, a = o - di * PPDE_Q;
                        float acc = 0.f;
#pragma unroll
                        for (int t = 0; t < 5; ++t) {
                            const int pp = di - t;
                            if (pp >= 0 && pp < nv) acc += sY[(t * PPDE_Q + a) * YS + pp];
                        }
                        const int i = p0 + cg * 32 + di;
                        if (i < L) sGc[i * PPDE_Q + a] += acc;
                    }
                    named_bar(2, NT_EPI);
                }
                if (tn == prm.tiles_per_chain - 1) {
                    // flush the chain's partial gradient (streaming float4 stores) and clear the accumulator
                    float4* dst = reinterpret_cast<float4*>(prm.Gc + ((size_t)k * prm.n + b) * NE);
                    float4* src = reinterpret_cast<float4*>(sGc);
                    for (int e = tid; e < NE / 4; e += NT_EPI) {
                        __stcs(dst + e, src[e]);
                        src[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    named_bar(2, NT_EPI);
                }
            }
        } else if (warp == WARP_MMA) {
            if (lane == 0) {
                const uint32_t idesc = make_idesc(128, prm.n_tile);
                const uint32_t ring_addr = smem_u32(ring);
                const int last_ksteps = (prm.kpad - (prm.nch - 1) * KCH) / 16;
                for (int it = 0; it < ntiles; ++it) {
                    const int gt = git + it;
                    const int buf = gt & 1;
                    if (gt >= 2) mbar_wait(&dempty[buf], (uint32_t)(((gt >> 1) + 1) & 1));
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + D_COL0 + buf * 128;
                    for (int kc = 0; kc < prm.nch; ++kc) {
                        mbar_wait(&full[m_slot], m_phase);
                        tc_fence_after();
                        const uint64_t dhi = make_b_desc(ring_addr + m_slot * SLOT_BYTES);
                        const uint64_t dlo = make_b_desc(ring_addr + m_slot * SLOT_BYTES + MAT_BYTES);
                        const int ksteps = (kc == prm.nch - 1) ? last_ksteps : KCH / 16;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint32_t a_hi = tmem_base + kc * (KCH / 2) + ks * 8;
                            const uint32_t a_lo = a_hi + prm.kpad / 2;
                            const uint64_t koff = (uint64_t)(ks * 2);
                            mma_ts(d_tmem, a_hi, dhi + koff, idesc, (kc | ks) ? 1u : 0u);
                            mma_ts(d_tmem, a_hi, dlo + koff, idesc, 1u);
                            mma_ts(d_tmem, a_lo, dhi + koff, idesc, 1u);
                        }
                        tc_commit(&empty[m_slot]);
                        if (++m_slot == BW_NSLOT) { m_slot = 0; m_phase ^= 1; }
                    }
                    tc_commit(&dfull[buf]);
                }
            }
        } else {
            // ===== PRODUCERS: adjoint rows =====
            const int ptid = threadIdx.x - (WARP_MMA + 1) * 32;           // 0..511
            const int pw = warp - 5;                                      // 0..15
            const int g = lane & 7, q = lane >> 3;
            const int rsub = 16 * (pw >> 2) + (pw & 3) + 4 * q;           // row inside a 64-row pass
            const float adj_scale = net.adj_scale;
            for (int it = 0; it < ntiles; ++it) {
                const int b = b_lo + it / prm.tiles_per_chain;
                const int tn = it - (it / prm.tiles_per_chain) * prm.tiles_per_chain;
                const int p0 = tn * prm.n_tile;
                const int valid = min(prm.n_tile, P - p0);
                if (tn == 0) {
                    // bucket this chain's winners by position (counting sort, then ascending channel order per bucket)
                    named_bar(1, BW_NT_PROD);
                    for (int i = ptid; i <= P; i += BW_NT_PROD) { sStart[i] = 0; if (i < P) sFill[i] = 0; }
                    named_bar(1, BW_NT_PROD);
                    const unsigned long long* keys = prm.mkey + ((size_t)b * prm.m.n_nets + k) * J2;
                    for (int j = ptid; j < J2; j += BW_NT_PROD) {
                        const unsigned long long key = keys[j];
                        const float mj = __uint_as_float((unsigned)(key >> 32));
                        const int pst = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu));
                        const bool active = (mj > 0.f) && pst >= 0 && pst < P;     // relu'(0) = 0
                        sPst[j] = active ? pst : -1;
                        sDj[j] = net.d[j];
                        if (active) atomicAdd(&sStart[pst + 1], 1);
                    }
                    named_bar(1, BW_NT_PROD);
                    if (pw == 0) {                                        // scan of P+1 counters by one warp
                        const int per = (P + 1 + 31) / 32;
                        int run = 0;
                        for (int i = lane * per; i < min((lane + 1) * per, P + 1); ++i) run += sStart[i];
                        int incl = run;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const int v = __shfl_up_sync(0xffffffffu, incl, o);
                            if (lane >= o) incl += v;
                        }
                        int base = incl - run;
                        for (int i = lane * per; i < min((lane + 1) * per, P + 1); ++i) { base += sStart[i]; sStart[i] = base; }
                    }
                    named_bar(1, BW_NT_PROD);
                    for (int j = ptid; j < J2; j += BW_NT_PROD) {
                        const int pst = sPst[j];
                        if (pst >= 0) sList[sStart[pst] + atomicAdd(&sFill[pst], 1)] = j;
                    }
                    named_bar(1, BW_NT_PROD);
                    for (int pp = ptid; pp < P; pp += BW_NT_PROD) {
                        const int s0 = sStart[pp], s1 = sStart[pp + 1];
                        for (int u = s0 + 1; u < s1; ++u) {
                            const int v = sList[u];
                            int w = u - 1;
                            while (w >= s0 && sList[w] > v) { sList[w + 1] = sList[w]; --w; }
                            sList[w + 1] = v;
                        }
                    }
                    named_bar(1, BW_NT_PROD);
                }
                int ls0[2], ls1[2];
                const uint8_t* mrow[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = 64 * i + rsub;
                    const bool ok = r < valid;
                    ls0[i] = ok ? sStart[p0 + r] : 0;
                    ls1[i] = ok ? sStart[p0 + r + 1] : 0;
                    mrow[i] = prm.r1mask + (((size_t)b * prm.m.n_nets + k) * P + p0 + (ok ? r : 0)) * 32;
                }
                for (int kc = 0; kc < prm.nch; ++kc) {
                    mbar_wait(&empty[p_slot], p_phase ^ 1);
                    unsigned char* mat_hi = ring + p_slot * SLOT_BYTES;
                    unsigned char* mat_lo = mat_hi + MAT_BYTES;
                    const int cb = kc * KCH;
                    const bool in0 = cb + 4 * g < prm.kpad, in1 = cb + 32 + 4 * g < prm.kpad;
                    // relu-mask bytes of my rows for this chunk (8 bytes = 64 channels), issued with the gathers
                    uint2 mb[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        mb[i] = make_uint2(0u, 0u);
                        if (ls1[i] > ls0[i]) mb[i] = __ldg(reinterpret_cast<const uint2*>(mrow[i] + (cb >> 3)));
                    }
                    // gather-sum the winners' W1 row segments (ascending channel order per row: deterministic),
                    // two winners of both rows per iteration -> up to 8 independent 16-byte L2 loads in flight per thread
                    float4 s0[2], s1[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) { s0[i] = make_float4(0.f, 0.f, 0.f, 0.f); s1[i] = s0[i]; }
                    const int maxlen = max(ls1[0] - ls0[0], ls1[1] - ls0[1]);
                    for (int u = 0; u < maxlen; u += 2) {
                        float dj[2][2];
                        float4 w0[2][2], w1[2][2];
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int v = 0; v < 2; ++v) {
                                dj[i][v] = 0.f;
                                w0[i][v] = make_float4(0.f, 0.f, 0.f, 0.f);
                                w1[i][v] = w0[i][v];
                                if (ls0[i] + u + v < ls1[i]) {
                                    const int jn = sList[ls0[i] + u + v];
                                    dj[i][v] = sDj[jn];
                                    const float* wrow = net.W1p + (size_t)jn * prm.kpad + cb;
                                    if (in0) w0[i][v] = __ldg(reinterpret_cast<const float4*>(wrow + 4 * g));
                                    if (in1) w1[i][v] = __ldg(reinterpret_cast<const float4*>(wrow + 32 + 4 * g));
                                }
                            }
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int v = 0; v < 2; ++v) {
                                s0[i].x = fmaf(dj[i][v], w0[i][v].x, s0[i].x); s0[i].y = fmaf(dj[i][v], w0[i][v].y, s0[i].y);
                                s0[i].z = fmaf(dj[i][v], w0[i][v].z, s0[i].z); s0[i].w = fmaf(dj[i][v], w0[i][v].w, s0[i].w);
                                s1[i].x = fmaf(dj[i][v], w1[i][v].x, s1[i].x); s1[i].y = fmaf(dj[i][v], w1[i][v].y, s1[i].y);
                                s1[i].z = fmaf(dj[i][v], w1[i][v].z, s1[i].z); s1[i].w = fmaf(dj[i][v], w1[i][v].w, s1[i].w);
                            }
                    }
                    // mask, scale, fp16 hi/lo split, swizzled store (zeros for rows without winners)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int r = 64 * i + rsub;
                        if (r >= prm.n_tile) continue;
                        const uint32_t w_lo = (g >> 1) < 4 ? mb[i].x : 0u;      // bytes 0..3 hold channels cb .. cb+31
                        const uint32_t nib0 = (w_lo >> (8 * (g >> 1) + 4 * (g & 1))) & 15u;          // channels cb + 4g ..
                        const uint32_t nib1 = (mb[i].y >> (8 * (g >> 1) + 4 * (g & 1))) & 15u;       // channels cb + 32 + 4g ..
                        const uint32_t msk = nib0 | (nib1 << 4);
                        const float ss[8] = {s0[i].x, s0[i].y, s0[i].z, s0[i].w, s1[i].x, s1[i].y, s1[i].z, s1[i].w};
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float x0 = ((msk >> (2 * e)) & 1u) ? ss[2 * e] * adj_scale : 0.f;
                            const float x1 = ((msk >> (2 * e + 1)) & 1u) ? ss[2 * e + 1] * adj_scale : 0.f;
                            const float h0 = h_round(x0), h1 = h_round(x1);
                            hi[e] = pack_h2(h0, h1);
                            lo[e] = pack_h2(x0 - h0, x1 - h1);
                        }
                        const int rbase = (r >> 3) * 1024 + (r & 7) * 128;
                        const int o0 = rbase + (((g >> 1) ^ (r & 7)) << 4) + ((g & 1) << 3);
                        const int o1 = rbase + ((((g >> 1) + 4) ^ (r & 7)) << 4) + ((g & 1) << 3);
                        *reinterpret_cast<uint2*>(mat_hi + o0) = make_uint2(hi[0], hi[1]);
                        *reinterpret_cast<uint2*>(mat_hi + o1) = make_uint2(hi[2], hi[3]);
                        *reinterpret_cast<uint2*>(mat_lo + o0) = make_uint2(lo[0], lo[1]);
                        *reinterpret_cast<uint2*>(mat_lo + o1) = make_uint2(lo[2], lo[3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full[p_slot]);
                    if (++p_slot == BW_NSLOT) { p_slot = 0; p_phase ^= 1; }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    if (warp == WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}


__global__ void cnn_grad_combine_kernel(int n, int NE, int n_nets, float scale, ppde_potts_t pm,
                                        const float* __restrict__ Gc, const float* __restrict__ Gp, int64_t Gp_stride,
                                        const int32_t* __restrict__ gp_rows, float* __restrict__ G, int64_t G_stride,
                                        const int32_t* __restrict__ g_rows) {
    const int b = blockIdx.x;
    const int wlo = pm.win_lo * PPDE_Q, whi = (pm.win_lo + pm.Lp) * PPDE_Q;     // multiples of 4
    float4* g = reinterpret_cast<float4*>(G + (int64_t)(g_rows ? g_rows[b] : b) * G_stride);
    const float* gp = Gp ? Gp + (int64_t)(gp_rows ? gp_rows[b] : b) * Gp_stride : nullptr;
    for (int q = threadIdx.x; q < NE / 4; q += blockDim.x) {
        float4 acc = __ldcs(reinterpret_cast<const float4*>(Gc + (size_t)b * NE) + q);
        for (int k = 1; k < n_nets; ++k) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(Gc + ((size_t)k * n + b) * NE) + q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float4 base = make_float4(0.f, 0.f, 0.f, 0.f);
        const int e = q * 4;
        if (gp && e >= wlo && e < whi) base = *reinterpret_cast<const float4*>(gp + (e - wlo));
        g[q] = make_float4(fmaf(scale, acc.x, base.x), fmaf(scale, acc.y, base.y), fmaf(scale, acc.z, base.z),
                           fmaf(scale, acc.w, base.w));
    }
}

}  // namespace tc
}  // namespace ppde

using namespace ppde;

// Best positions-per-tile: multiple of 16 in [64,128] minimising padded work.
static int choose_n_tile(int P, int* tiles) {
    int best_n = 128, best_cost = 1 << 30;
    for (int nt = 128; nt >= 64; nt -= 16) {
        const int t = (P + nt - 1) / nt;
        const int cost = t * nt + 8 * t;          // padded positions + a small per-tile overhead
        if (cost < best_cost) { best_cost = cost; best_n = nt; *tiles = t; }
    }
    return best_n;
}

// positions-per-tile for the 2-CTA kernel: multiple of 32 (N of a cta_group::2 MMA), split in two halves
static int choose_n_tile2(int P, int* tiles) {
    int best_n = 128, best_cost = 1 << 30;
    for (int nt = 128; nt >= 64; nt -= 32) {
        const int t = (P + nt - 1) / nt;
        const int cost = t * nt + 8 * t;
        if (cost < best_cost) { best_cost = cost; best_n = nt; *tiles = t; }
    }
    return best_n;
}

static int g_forward_variant = -1;         // -1: read PPDE_TC_CTAS once (default 2); 1 or 2
extern "C" int ppde_set_forward_variant(int ctas) { g_forward_variant = (ctas == 1) ? 1 : 2; return 0; }

extern "C" int ppde_cnn_forward_tc(const ppde_cnn_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n,
                                   unsigned long long* mkey, uint8_t* r1mask, void* stream) {
    if (n <= 0) return 0;
    if (m->C > 256 || m->P < 1) return (int)cudaErrorInvalidValue;       // A must fit 256 TMEM columns
    if (g_forward_variant < 0) {
        const char* e = getenv("PPDE_TC_CTAS");
        g_forward_variant = (e && e[0] == '1') ? 1 : 2;
    }
    tc::Params prm;
    prm.m = *m;
    prm.aa = aa;
    prm.aa_stride = aa_stride;
    prm.n = n;
    prm.mkey = mkey;
    prm.r1mask = r1mask;
    prm.kpad = (m->C + 15) / 16 * 16;
    prm.nch = (prm.kpad + tc::KCH - 1) / tc::KCH;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int MT = (2 * m->C + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (g_forward_variant == 2) {
        prm.n_tile = choose_n_tile2(m->P, &prm.tiles_per_chain);
        prm.MT = (MT + 1) / 2;                                             // channel-tile pairs
        const int combos = m->n_nets * prm.MT;
        prm.ctas_per_combo = (sms / 2) / combos;                           // cluster pairs per combo
        if (prm.ctas_per_combo < 1) prm.ctas_per_combo = 1;
        if (prm.ctas_per_combo > n) prm.ctas_per_combo = n;
        const size_t smem = (size_t)tc::NSLOT2 * tc::SLOT2_BYTES + (size_t)100 * prm.nch * tc::KCH * sizeof(float) +
                            32 * sizeof(uint64_t) + 1024;
        static size_t configured2 = 0;
        if (smem > configured2) {
            cudaError_t e = cudaFuncSetAttribute(tc::cnn_forward_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return (int)e;
            configured2 = smem;
        }
        tc::cnn_forward_tc2_kernel<<<2 * combos * prm.ctas_per_combo, tc::NTHREADS, smem, st>>>(prm);
        return launch_done();
    }
    prm.n_tile = choose_n_tile(m->P, &prm.tiles_per_chain);
    prm.MT = MT;
    const int combos = m->n_nets * prm.MT;
    prm.ctas_per_combo = sms / combos;
    if (prm.ctas_per_combo < 1) prm.ctas_per_combo = 1;
    if (prm.ctas_per_combo > n) prm.ctas_per_combo = n;
    const size_t smem = (size_t)tc::NSLOT * tc::SLOT_BYTES + (size_t)100 * prm.nch * tc::KCH * sizeof(float) +
                        16 * sizeof(uint64_t) + 1024;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::cnn_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    tc::cnn_forward_tc_kernel<<<combos * prm.ctas_per_combo, tc::NTHREADS, smem, st>>>(prm);
    return launch_done();
}

extern "C" int ppde_cnn_backward_tc(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa, int32_t aa_stride,
                                    int32_t n, const unsigned long long* mkey, float lamda,
                                    const float* Gp, int64_t Gp_stride, const int32_t* gp_rows,
                                    float* G, int64_t G_stride, const int32_t* g_rows, const uint8_t* r1mask,
                                    float* scratch, void* stream) {
    if (n <= 0) return 0;
    if (m->C > 256 || m->P < 1 || !scratch || !r1mask) return (int)cudaErrorInvalidValue;
    tc::BwdParams prm;
    prm.m = *m; prm.pm = *pm; prm.aa = aa; prm.aa_stride = aa_stride; prm.n = n; prm.mkey = mkey;
    prm.Gc = scratch;
    prm.r1mask = r1mask;
    prm.n_tile = choose_n_tile(m->P, &prm.tiles_per_chain);
    prm.kpad = (m->C + 15) / 16 * 16;
    prm.nch = (prm.kpad + tc::KCH - 1) / tc::KCH;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    prm.ctas_per_net = sms / m->n_nets;
    if (prm.ctas_per_net < 1) prm.ctas_per_net = 1;
    if (prm.ctas_per_net > n) prm.ctas_per_net = n;
    const int C = m->C, P = m->P, L = m->L, J2 = 2 * C;
    (void)C;
    const size_t smem = 1024 + (size_t)tc::BW_NSLOT * tc::SLOT_BYTES +
                        ((size_t)L * PPDE_Q + 100 * tc::YS + J2) * sizeof(float) +
                        ((size_t)J2 + (P + 1) + P + J2) * sizeof(int) + 8 + 16 * sizeof(uint64_t);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::cnn_backward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured = smem;
    }
    cudaStream_t st = (cudaStream_t)stream;
    tc::cnn_backward_tc_kernel<<<m->n_nets * prm.ctas_per_net, tc::BW_NTHREADS, smem, st>>>(prm);
    int r = launch_done();
    if (r) return r;
    tc::cnn_grad_combine_kernel<<<n, 256, 0, st>>>(n, L * PPDE_Q, m->n_nets, lamda / (float)m->n_nets, *pm, scratch,
                                                   Gp, Gp_stride, gp_rows, G, G_stride, g_rows);
    return launch_done();
}
