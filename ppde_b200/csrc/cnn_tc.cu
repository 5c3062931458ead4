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
#include "tc_common.cuh"
#include <cuda_fp16.h>
#include <cstdlib>

namespace ppde {
namespace tc {

constexpr int NT_EPI = 128;
constexpr int NT_PROD = 512;                       // 16 producer warps (4 per scheduler): the producers are issue/latency bound
constexpr int NTHREADS = NT_EPI + 32 + NT_PROD;   // 672
constexpr int WARP_MMA = 4;
constexpr int NSLOT = 3;
constexpr int MAT_BYTES = 128 * KCH * 2;           // one [128 x 64] fp16 operand matrix (16 KB)
constexpr int SLOT_BYTES = 2 * MAT_BYTES;          // hi + lo
constexpr int D_COL0 = 256;                        // accumulators: columns [256,384) and [384,512)

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
    long long* prof;      // optional [grid][16] cycle counters (PPDE_TC_PROFILE=1 builds of the 2-CTA kernel only)
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
// r1 production cost per MMA flop halves, and so does the shared-memory traffic of the B operand per CTA.
//
// Tiles: tile t of a chain covers positions [128 t, 128 t + N_t), N_t = min(128, roundup16(P - 128 t)); CTA `rank`
// produces rows [rank N_t/2, (rank+1) N_t/2).  Rows past P-1 replicate position P-1: a duplicate column can never beat
// the original under the strict `>` / lowest-index rule, so the epilogue needs no per-column validity test.
//
// Synchronisation (per 64-wide K chunk = ring slot):
//   producers (16 warps / CTA)  --fullL[slot] (local, 16 warp arrivals)-->  leader's MMA thread (rank 0)
//                                                                     \-->  rank 1's forwarder thread --fullR[slot]
//                                                                           (one remote arrive / chunk)--> leader
//   leader MMA  --tcgen05.commit multicast--> empty[slot] in both CTAs (slot free), dfull[buf] in both CTAs
//   epilogue warps (4 / CTA) --dempty[buf] in the leader (one arrive per warp)--> leader MMA
// The only cluster-scope (expensive) arrives are 1 per chunk (forwarder) and 4 per tile (rank 1's epilogue warps).
constexpr int NSLOT2 = 6;
constexpr int MAT2_BYTES = 64 * KCH * 2;          // one [64 x 64] fp16 operand half-matrix (8 KB)
constexpr int SLOT2_BYTES = 2 * MAT2_BYTES;       // hi + lo
constexpr int NT2 = 128;                          // positions per full tile
constexpr int WARP_MMA2 = 20;                     // MMA issuer = highest warp id of its scheduler (top arbitration priority)


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
// explicit shared-space accesses with 32-bit addresses and immediate offsets (generic LD/ST cost 64-bit address math
// and the slower generic path; the table pointers are data-dependent, so the compiler cannot infer the space)
template <int IMM>
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0+%1], {%2, %3};" ::"r"(addr), "n"(IMM), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    float2 r;
    asm("sub.rn.f32x2 %0, %1, %2;"
        : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return r;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// one 64-wide K chunk of one r1 row: 8 channels per thread (4g..4g+3 and 32+4g..+3 of the chunk)
template <int KC>
__device__ __forceinline__ void produce_chunk(const uint32_t (&ra)[5], uint32_t sa0, uint32_t sa1, bool emit_mask,
                                              uint32_t& mbits) {
    constexpr int CB = KC * KCH * 4;               // byte offset of the chunk inside a table row
    const float4 z0 = lds128<CB>(ra[0]);
    const float4 z1 = lds128<CB + 128>(ra[0]);
    float2 a0 = make_float2(z0.x, z0.y), a1 = make_float2(z0.z, z0.w);
    float2 a2 = make_float2(z1.x, z1.y), a3 = make_float2(z1.z, z1.w);
#pragma unroll
    for (int t = 1; t < 5; ++t) {
        const float4 u0 = lds128<CB>(ra[t]);
        const float4 u1 = lds128<CB + 128>(ra[t]);
        a0 = add2(a0, make_float2(u0.x, u0.y)); a1 = add2(a1, make_float2(u0.z, u0.w));
        a2 = add2(a2, make_float2(u1.x, u1.y)); a3 = add2(a3, make_float2(u1.z, u1.w));
    }
    if (emit_mask) {                               // warp-uniform
        const float v[8] = {a0.x, a0.y, a1.x, a1.y, a2.x, a2.y, a3.x, a3.y};
        uint32_t nib = 0u;
#pragma unroll
        for (int e = 0; e < 8; ++e) nib |= (__float_as_int(v[e]) > 0 ? 1u : 0u) << e;
        mbits |= nib << (8 * KC);
    }
    float2 x[4] = {a0, a1, a2, a3};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        x[e].x = fmaxf(x[e].x, 0.f); x[e].y = fmaxf(x[e].y, 0.f);
        const float2 h = make_float2(h_trunc(x[e].x), h_trunc(x[e].y));
        const float2 l = sub2(x[e], h);
        hi[e] = pack_h2(h.x, h.y);
        lo[e] = pack_h2(l.x, l.y);
    }
    sts64<0>(sa0, hi[0], hi[1]);
    sts64<0>(sa1, hi[2], hi[3]);
    sts64<MAT2_BYTES>(sa0, lo[0], lo[1]);
    sts64<MAT2_BYTES>(sa1, lo[2], lo[3]);
}

template <int NCH, bool PROF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
cnn_forward_tc2_kernel(const __grid_constant__ Params prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int C = prm.m.C, P = prm.m.P, J2 = 2 * C;
    constexpr int KS = NCH * KCH;
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* sT0 = reinterpret_cast<float*>(ring + NSLOT2 * SLOT2_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT0 + 100 * KS);
    uint64_t* fullL = bars;                // [NSLOT2] local: this CTA's producer warps
    uint64_t* fullR = fullL + NSLOT2;      // [NSLOT2] leader's copy is live: forwarded "rank 1 is full"
    uint64_t* empty = fullR + NSLOT2;      // [NSLOT2] local (multicast commit)
    uint64_t* dfull = empty + NSLOT2;      // [2] local (multicast commit)
    uint64_t* dempty = dfull + 2;          // [2] leader's copy is live
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
    const int tpc = prm.tiles_per_chain;
    const int ntiles = (b_hi - b_lo) * tpc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_last = ((P - (tpc - 1) * NT2) + 15) & ~15;   // N of a chain's last tile (multiple of 16, <= 128)

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
        for (int s = 0; s < NSLOT2; ++s) { mbar_init(&fullL[s], NT_PROD / 32); mbar_init(&fullR[s], 1); mbar_init(&empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&dfull[d], 1); mbar_init(&dempty[d], 2 * (NT_EPI / 32)); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA2) {
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
        // running (max, first arg-max) of the RAW accumulator: v = relu(u * unscale + b1) is monotone in u (unscale > 0)
        const int j = mt * 128 + warp * 32 + lane;
        const float bias = (j < J2) ? net.b1[j] : 0.f;
        const float unscale = 1.f / (net.w1_scale * net.r1_scale);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL0;
        float best = -3.0e38f;
        int bp = 0;
        int b = b_lo, tn = 0;
        long long pc[3] = {0, 0, 0};
        long long tp = PROF ? clock64() : 0;
        for (int it = 0; it < ntiles; ++it) {
            const int buf = it & 1;
            const int p0 = tn * NT2;
            const int nt = (tn == tpc - 1) ? n_last : NT2;
            if (tn == 0) { best = -3.0e38f; bp = 0; }
            int bl = bp - p0;                                  // arg-max relative to this tile (immediates below)
            mbar_wait(&dfull[buf], (uint32_t)((it >> 1) & 1));
            if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
            tc_fence_after();
            const uint32_t ta = lane_addr + buf * 128;
#pragma unroll
            for (int h = 0; h < 4; ++h) {                      // 32 columns at a time, two x16 loads in flight
                if (h * 32 < nt) {
                    uint32_t r0[16], r1[16];
                    tmem_ld16(ta + h * 32, r0);
                    const bool second = h * 32 + 16 < nt;
                    if (second) tmem_ld16(ta + h * 32 + 16, r1);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float u = __uint_as_float(r0[i]);
                        if (u > best) { best = u; bl = h * 32 + i; }
                    }
                    if (second) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float u = __uint_as_float(r1[i]);
                            if (u > best) { best = u; bl = h * 32 + 16 + i; }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&dempty[buf]); else mbar_arrive_cluster(&dempty[buf], 0);
            }
            if (PROF) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
            bp = bl + p0;
            if (tn == tpc - 1) {
                if (j < J2) {
                    float v = fmaf(best, unscale, bias);
                    int pp = bp;
                    if (!(v > 0.f)) { v = 0.f; pp = 0; }       // relu; all-nonpositive column -> (0, position 0)
                    const unsigned long long key =
                        ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)pp);
                    prm.mkey[((size_t)b * prm.m.n_nets + k) * J2 + j] = key;
                }
                tn = 0; ++b;
            } else {
                ++tn;
            }
        }
        if (PROF && threadIdx.x == 0 && prm.prof) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; }
    } else if (warp == WARP_MMA2) {
        if (rank == 0) {
            // ===== MMA ISSUER (leader CTA): the whole warp runs the loop warp-uniformly (descriptors live in uniform
            // registers, no per-MMA broadcasts); one elected lane issues the MMAs and the commits =====
            const uint32_t idesc_full = make_idesc(256, NT2), idesc_last = make_idesc(256, n_last);
            const uint32_t ring_addr = smem_u32(ring);
            int slot = 0, tn = 0;
            uint32_t sphase = 0;
            const int last_ksteps = (prm.kpad - (NCH - 1) * KCH) / 16;
            const uint32_t a_lo_off = (uint32_t)(prm.kpad / 2);
            long long pc[4] = {0, 0, 0, 0};
            long long tp = PROF ? clock64() : 0;
            const long long tstart = tp;
            for (int it = 0; it < ntiles; ++it) {
                const int buf = it & 1;
                const uint32_t idesc = (tn == tpc - 1) ? idesc_last : idesc_full;
                if (++tn == tpc) tn = 0;
                if (it >= 2) mbar_wait_cluster(&dempty[buf], (uint32_t)(((it >> 1) + 1) & 1));
                if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + D_COL0 + buf * 128;
#pragma unroll 1
                for (int kc = 0; kc < NCH; ++kc) {
                    mbar_wait(&fullL[slot], sphase);
                    if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                    mbar_wait_cluster(&fullR[slot], sphase);
                    if (PROF) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
                    tc_fence_after();
                    const uint64_t dhi = make_b_desc(ring_addr + slot * SLOT2_BYTES);
                    const uint64_t dlo = make_b_desc(ring_addr + slot * SLOT2_BYTES + MAT2_BYTES);
                    const int ksteps = (kc == NCH - 1) ? last_ksteps : KCH / 16;
                    const uint32_t a_hi0 = tmem_base + kc * (KCH / 2);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < KCH / 16; ++ks) {
                            if (ks < ksteps) {
                                const uint32_t a_hi = a_hi0 + ks * 8;
                                mma_ts2(d_tmem, a_hi, dhi + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
                                mma_ts2(d_tmem, a_hi, dlo + (uint64_t)(ks * 2), idesc, 1u);
                                mma_ts2(d_tmem, a_hi + a_lo_off, dhi + (uint64_t)(ks * 2), idesc, 1u);
                            }
                        }
                        tc_commit2(&empty[slot]);
                        if (kc == NCH - 1) tc_commit2(&dfull[buf]);
                    }
                    __syncwarp();
                    if (PROF) { const long long t1 = clock64(); pc[3] += t1 - tp; tp = t1; }
                    if (++slot == NSLOT2) { slot = 0; sphase ^= 1; }
                }
            }
            if (PROF && prm.prof && lane == 0) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[3] = pc[0]; o[4] = pc[1]; o[5] = pc[2]; o[6] = pc[3]; o[7] = clock64() - tstart; }
        } else if (lane == 0) {
            // ===== FORWARDER (rank 1): one remote arrive per chunk instead of one per producer warp =====
            int slot = 0;
            uint32_t sphase = 0;
            long long pc[2] = {0, 0};
            long long tp = PROF ? clock64() : 0;
            for (int c = 0; c < ntiles * NCH; ++c) {
                mbar_wait(&fullL[slot], sphase);
                if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
                mbar_arrive_cluster(&fullR[slot], 0);
                if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                if (++slot == NSLOT2) { slot = 0; sphase ^= 1; }
            }
            if (PROF && prm.prof) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[3] = pc[0]; o[4] = pc[1]; }
        }
    } else {
        // ===== PRODUCERS (both CTAs): one row per thread, rows [rank*N_t/2, (rank+1)*N_t/2) of every tile =====
        // lane = g + 8q: g = channel group (conflict-free 128-byte LDS phases), q = one of the warp's 4 rows;
        // rows {x, x+4, x+8, x+12} per warp keep the 8-byte swizzled stores conflict-free.
        const int pw = warp - 4;                               // producers are warps 4..19
        const int g = lane & 7, q = lane >> 3;
        const int r = 16 * (pw >> 2) + (pw & 3) + 4 * q;      // local row 0..63
        const uint32_t t0addr = smem_u32(sT0) + 16 * g;
        // element (row r, k) at (r/8)*1024 + (r%8)*128 + ((k/8) ^ (r%8))*16 + (k%8)*2;  k = 4g  and  k = 32 + 4g
        const uint32_t o0 = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((g >> 1)) ^ (r & 7)) << 4) + ((g & 1) << 3));
        const uint32_t st0 = smem_u32(ring) + o0, st1 = smem_u32(ring) + (o0 ^ 64u);
        int slot = 0;
        uint32_t phase = 0;
        int b = b_lo, tn = 0;
        uint32_t an[5] = {0u, 0u, 0u, 0u, 0u};               // residues of MY row's 5 taps, one tile ahead
        auto row_pos = [&](int tt) {
            const int nt = (tt == tpc - 1) ? n_last : NT2;
            return tt * NT2 + (int)rank * (nt >> 1) + r;
        };
        auto load_aa = [&](int bb, int tt) {
            const int pos = min(row_pos(tt), P - 1);          // rows past the end replicate the last position
            const uint8_t* ap = prm.aa + (size_t)bb * prm.aa_stride + pos;
#pragma unroll
            for (int t = 0; t < 5; ++t) an[t] = ap[t];
        };
        if (ntiles > 0) load_aa(b, 0);
        int emit_ctr = 0;                                     // the MP clusters sharing a chain block take turns with the mask
        long long pc[3] = {0, 0, 0};
        long long tp = PROF ? clock64() : 0;
        for (int it = 0; it < ntiles; ++it) {
            const int nt = (tn == tpc - 1) ? n_last : NT2;
            const bool active = r < (nt >> 1);
            const int pos = row_pos(tn);
            uint32_t ra[5];
#pragma unroll
            for (int t = 0; t < 5; ++t) ra[t] = t0addr + (uint32_t)((t * PPDE_Q + (int)an[t]) * (KS * 4));
            const int bcur = b;
            if (++tn == tpc) { tn = 0; ++b; }
            if (it + 1 < ntiles) load_aa(b, tn);              // in flight during this tile's chunks
            const bool emit_mask = (prm.r1mask != nullptr) && (emit_ctr == mp);
            if (++emit_ctr == MP) emit_ctr = 0;
            uint32_t mbits = 0u;
#pragma unroll
            for (int kc = 0; kc < NCH; ++kc) {
                mbar_wait(&empty[slot], phase ^ 1);
                if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
                if (active) {
                    const uint32_t so = (uint32_t)(slot * SLOT2_BYTES);
                    if (kc == 0) produce_chunk<0>(ra, st0 + so, st1 + so, emit_mask, mbits);
                    if (kc == 1) produce_chunk<1>(ra, st0 + so, st1 + so, emit_mask, mbits);
                    if (kc == 2) produce_chunk<2>(ra, st0 + so, st1 + so, emit_mask, mbits);
                    if (kc == 3) produce_chunk<3>(ra, st0 + so, st1 + so, emit_mask, mbits);
                }
                if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&fullL[slot]);
                if (PROF) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
                if (++slot == NSLOT2) { slot = 0; phase ^= 1; }
            }
            if (emit_mask) {
                // word w of a row's 256 mask bits = channels 32w..32w+31 = nibble w of the row's 8 lanes:
                // 8x8 nibble transpose across those lanes (3 butterfly stages), then one coalesced 32-byte store per row
                uint32_t word = mbits;
                uint32_t o = __shfl_xor_sync(0xffffffffu, word, 4);
                word = (g & 4) ? ((word & 0xFFFF0000u) | ((o >> 16) & 0x0000FFFFu)) : ((word & 0x0000FFFFu) | ((o << 16) & 0xFFFF0000u));
                o = __shfl_xor_sync(0xffffffffu, word, 2);
                word = (g & 2) ? ((word & 0xFF00FF00u) | ((o >> 8) & 0x00FF00FFu)) : ((word & 0x00FF00FFu) | ((o << 8) & 0xFF00FF00u));
                o = __shfl_xor_sync(0xffffffffu, word, 1);
                word = (g & 1) ? ((word & 0xF0F0F0F0u) | ((o >> 4) & 0x0F0F0F0Fu)) : ((word & 0x0F0F0F0Fu) | ((o << 4) & 0xF0F0F0F0u));
                if (active && pos < P)
                    reinterpret_cast<uint32_t*>(prm.r1mask + (((size_t)bcur * prm.m.n_nets + k) * P + pos) * 32)[g] = word;
            }
        }
        if (PROF && prm.prof && lane == 0 && (pw == 0 || pw == 15)) {
            long long* o = prm.prof + (size_t)blockIdx.x * 16 + (pw == 0 ? 8 : 11); o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2];
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                    // the peer may still be reading my smem / signalling my barriers
    if (warp == WARP_MMA2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// =====================================================================================================
// INCREMENTAL forward (same 2-CTA tcgen05 skeleton as cnn_forward_tc2_kernel).
//
// A proposal y differs from the chain's current state x in at most S residues, and a residue change at i only moves
// the conv rows p in [i-4, i]: the max-pool of y over all P positions is re-used from x except for the few 16-position
// BLOCKS that contain such a row.  Per (pool row, net, block q, channel j) the pool `bkey` holds
//     key = ordered(raw accumulator max over the block) << 32 | (0xFFFFFFFF - first arg-max position)
// (ordered() = the usual order-preserving map of fp32 bits, -0 canonicalised), so the chain-level winner - first
// arg-max of the raw accumulator, the quantity cnn_forward_tc2_kernel tracks - is the 64-bit maximum over the blocks.
// The kernel recomputes only the dirty blocks of every chain (bit q of dmask[b]; dmask == NULL: every block).  The dirty
// blocks of a CTA's chains form one sequence (cnn_inc_scan_kernel: exclusive prefix `boff`, compact list `blist`), cut
// into MMA tiles of 8 blocks = 128 positions REGARDLESS of chain boundaries: a tcgen05.mma costs the same ~60 cycles at
// N = 32 as at N = 128, so one short tile per chain would be bound by the MMA count, not by the flops.  The epilogue
// copies the clean blocks' keys from the current row to the proposal row and writes the same `mkey` the full kernel
// writes, bit for bit: every accumulator element depends only on its own r1 row and W1 row, in the same K order.  The relu-mask rows of the dirty positions go to the proposal
// row of the r1mask pool (the clean rows were copied there by cnn_dirty_kernel).
// Block of the max-pool cache: PB conv-output positions (a tile of the incremental kernel = 128 positions = PB_TILE blocks).
// Round 1 used 16-position blocks; a changed residue dirties the rows [i-4, i], i.e. 1 + 4/PB blocks on average: 20 positions
// to recompute per mutation at PB = 16, 12 at PB = 8 (and twice as many, half as large, keys per pool row).
constexpr int PB_SHIFT = 3;
constexpr int PB = 1 << PB_SHIFT;
constexpr int PB_TILE = 128 / PB;              // blocks per MMA tile
constexpr int PB_MAXNB = 32;                   // blocks per row <= 32 (dirty-block masks are 32-bit words)
// Two epilogue groups (role counters, r02: the epilogue - 16 block arg-maxima and key stores per channel and tile, one warp per
// scheduler - was the critical path at 5.5 k cycles per tile with the MMA issuer waiting 3.8 k for a free accumulator):
// warps 0-3 take the even tiles (accumulator buffer 0), warps 24-27 the odd ones (buffer 1); both sets sit on the TMEM lane
// quarters 0-3 (warp id mod 4).  Warps 21-23 are idle padding.
constexpr int INC_NTHREADS = 28 * 32;
constexpr int INC_EPI2_WARP0 = 24;
struct IncParams {
    ppde_cnn_t m;
    const uint8_t* aa;                  // proposal states [n, aa_stride]
    int aa_stride;
    int n;
    unsigned long long* mkey;           // [n, n_nets, 2C]
    uint8_t* r1mask;                    // pool [rows, n_nets, P, 32] or NULL
    const uint32_t* dmask;              // [n] dirty-block bits, NULL = every block is dirty (full evaluation into the pool)
    unsigned long long* mkey_pool;      // optional pool [rows, n_nets, 2C][2]: the two largest RAW block keys of every pool row (see
                                        // cnn_inc_merge_kernel); read for the current row by the merge kernel, written for the proposal row
    unsigned long long* bkey;           // pool [rows, n_nets, NB, 2C]
    int32_t* btab;                      // pool [rows, NB]: the pool row whose slot holds block q of this row (keys AND relu-mask rows):
                                        // a proposal row only POINTS at the clean blocks of the current state instead of copying them
    const int32_t* rows_x;              // [n] pool row of the current state (NULL only with dmask == NULL)
    const int32_t* rows_y;              // [n] pool row of the proposal, NULL = row_base_y + b
    int row_base_y;
    const int32_t* boff;                // [n] dirty blocks of the CTA's chains before chain b (cnn_inc_scan_kernel)
    const int32_t* gtot;                // [ctas_per_combo] dirty blocks per chain range
    const uint32_t* blist;              // [n * PB_MAXNB] compact (chain << 5 | block) list, range r starts at PB_MAXNB * b_lo(r)
    int NB;                             // blocks of PB positions per chain = ceil(P / PB) <= PB_MAXNB
    int ctas_per_combo;
    int MT;                             // channel-tile pairs
    int kpad;
    long long* prof;                    // optional [grid][16] role-level cycle counters (ppde_set_forward_profile; tools/prof_inc.py)
    int dbg;                            // timing experiments only (PPDE_INC_DEBUG): 1 = no key traffic, 2 = no r1 production,
                                        // 4 = no residue loads (results are wrong with any bit set)
};

__device__ __forceinline__ int nth_set_bit(uint32_t mask, int k) {
    for (int i = 0; i < k; ++i) mask &= mask - 1u;
    return __ffs((int)mask) - 1;
}
__device__ __forceinline__ uint32_t f32_ordered(float u) {
    const uint32_t b = __float_as_uint(u + 0.0f);                  // -0 -> +0
    return b ^ ((uint32_t)((int32_t)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float f32_unordered(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
// chain-level winner key as the full forward kernel writes it (relu'd maximum << 32 | 0xFFFFFFFF - first arg-max) from the
// RAW winner of a pool row (ordered raw accumulator maximum << 32 | 0xFFFFFFFF - first arg-max)
__device__ __forceinline__ unsigned long long winner_from_raw(unsigned long long raw, float unscale, float bias) {
    const float u = f32_unordered((uint32_t)(raw >> 32));
    int pp = (int)(0xFFFFFFFFu - (uint32_t)(raw & 0xFFFFFFFFull));
    float v = fmaf(u, unscale, bias);
    if (!(v > 0.f)) { v = 0.f; pp = 0; }                    // relu; all-nonpositive column -> (0, position 0)
    return ((unsigned long long)__float_as_uint(v) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)pp);
}

// Cluster-scope barrier traffic WITHOUT release/acquire fences.  What these barriers order is never generic-proxy global
// data: (a) "rank 1's operand chunk is in ITS shared memory" - already performed there and fenced to the async proxy by the
// producers before the forwarder saw the local barrier flip; (b) "my tcgen05.ld of the accumulator has completed" -
// guaranteed by tcgen05.wait::ld + tcgen05.fence::before_thread_sync.  The default .release.cluster arrive compiles to
// MEMBAR.ALL.GPU + ERRBAR (and the acquire wait to CCTL.IVALL): ~600 cycles each and, once the thread has global stores
// in flight, a full write round trip - with one short tile per chain these would bound the incremental kernel.
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t cta) {
    uint32_t raddr;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_relaxed(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "WAIT_LOOP_R:\n"
        " mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE_R;\n"
        " bra WAIT_LOOP_R;\n"
        "DONE_R:\n"
        "}\n" ::"r"(a), "r"(parity)
        : "memory");
}

// (value, first arg-max) of the 8 raw accumulators r[o .. o+8) held as bit patterns
__device__ __forceinline__ void argmax8(const uint32_t (&r)[16], int o, float& best, int& bidx) {
    float v[4]; int ix[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = __uint_as_float(r[o + 2 * i]), b = __uint_as_float(r[o + 2 * i + 1]);
        const bool gt = b > a;
        v[i] = gt ? b : a; ix[i] = gt ? 2 * i + 1 : 2 * i;
    }
#pragma unroll
    for (int w = 2; w >= 1; w >>= 1) {
#pragma unroll
        for (int i = 0; i < w; ++i) {
            const bool gt = v[2 * i + 1] > v[2 * i];
            v[i] = gt ? v[2 * i + 1] : v[2 * i]; ix[i] = gt ? ix[2 * i + 1] : ix[2 * i];
        }
    }
    best = v[0]; bidx = ix[0];
}

template <int NCH>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(INC_NTHREADS, 1)
cnn_forward_inc_kernel(const __grid_constant__ IncParams prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int C = prm.m.C, P = prm.m.P, J2 = 2 * C, NB = prm.NB;
    constexpr int KS = NCH * KCH;
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* sT0 = reinterpret_cast<float*>(ring + NSLOT2 * SLOT2_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sT0 + 100 * KS);
    uint64_t* fullL = bars;
    uint64_t* fullR = fullL + NSLOT2;
    uint64_t* empty = fullR + NSLOT2;
    uint64_t* dfull = empty + NSLOT2;
    uint64_t* dempty = dfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int MP = prm.MT;
    const int combo = pair / prm.ctas_per_combo;
    const int within = pair - combo * prm.ctas_per_combo;
    const bool idle = combo >= prm.m.n_nets * MP;
    const int k = idle ? 0 : combo / MP, mp = idle ? 0 : combo - k * MP;
    const int mt = mp * 2 + (int)rank;
    const ppde_cnn_net_t net = prm.m.net[k];
    const int b_lo = idle ? 0 : (int)((int64_t)prm.n * within / prm.ctas_per_combo);
    const int b_hi = idle ? 0 : (int)((int64_t)prm.n * (within + 1) / prm.ctas_per_combo);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t all_blocks = (NB >= 32) ? 0xFFFFFFFFu : ((1u << NB) - 1u);
    auto ld_mask = [&](int b) -> uint32_t {
        return (b < b_hi) ? (prm.dmask ? (__ldg(prm.dmask + b) & all_blocks) : all_blocks) : 0u;
    };
    const int G = idle ? 0 : __ldg(prm.gtot + within);       // dirty blocks of my chains
    const int ntiles = (G + PB_TILE - 1) / PB_TILE;          // tiles of PB_TILE blocks (the last one may be shorter)
    const uint32_t* bl = prm.blist + (size_t)b_lo * PB_MAXNB;

    for (int e = threadIdx.x; e < 100 * KS; e += INC_NTHREADS) {
        const int row = e / KS, c = e - row * KS;
        float v = 0.f;
        if (c < C) {
            v = net.T0[(size_t)row * C + c];
            if (row < PPDE_Q) v += net.b0[c];
        }
        sT0[e] = v * net.r1_scale;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSLOT2; ++s) { mbar_init(&fullL[s], NT_PROD / 32); mbar_init(&fullR[s], 1); mbar_init(&empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&dfull[d], 1); mbar_init(&dempty[d], 2 * (NT_EPI / 32)); }
        fence_barrier_init();
    }
    if (warp == WARP_MMA2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {      // A = this CTA's W1 tile (fp16 hi/lo) -> tensor memory, as in the full kernel
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
    cluster_sync_all();
    tc_fence_after();

    if (warp < 4 || warp >= INC_EPI2_WARP0) {
        // ===== EPILOGUE (both CTAs, two groups of 4 warps: group g takes the tiles T = g, g + 2, ..): thread = channel j; per
        // TILE: the PB_TILE blocks' keys go straight to their final place in the proposal rows of the pool (no per-chain work
        // here: cnn_inc_merge_kernel forms mkey) =====
        const int eg = warp >= INC_EPI2_WARP0 ? 1 : 0, ew = warp & 3;
        const int j = mt * 128 + ew * 32 + lane;
        const bool jok = j < J2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(ew * 32) << 16) + D_COL0;
        const size_t row_keys = (size_t)prm.m.n_nets * NB * J2;            // keys per pool row
        const size_t koff = ((size_t)k * NB) * J2 + j;
        // lane l (mod PB_TILE) holds block l of the tile.  Three-stage prefetch, each stage one tile apart so that no load is waited
        // for in the iteration that issues it: entry (chain << 5 | block) three tiles ahead, the chain's pool rows two tiles ahead,
        // the block-table lookup that picks the slot one tile ahead (the table entry depends on the rows: a two-level chain that
        // cost ~700 cycles per tile when both levels were issued in the same iteration).
        auto ld_ent = [&](int T) -> uint32_t {
            const int g = PB_TILE * T + (lane & (PB_TILE - 1));
            return (g < G) ? __ldg(bl + g) : 0xFFFFFFFFu;                   // all ones: no block
        };
        auto ld_rows = [&](uint32_t ent, int& ry, int& rx) {
            ry = 0; rx = -1;
            if (ent == 0xFFFFFFFFu) return;
            const int b = (int)(ent >> 5);
            ry = prm.rows_y ? __ldg(prm.rows_y + b) : prm.row_base_y + b;
            if (prm.rows_x) rx = __ldg(prm.rows_x + b);
        };
        // slot that receives block (b, q): the proposal row's own, unless the current row still points at it (then the current
        // row's own slot is free: it is referenced by neither row) - see cnn_inc_merge_kernel for the table update
        auto ld_slot = [&](uint32_t ent, int ry, int rx) -> int {
            if (ent == 0xFFFFFFFFu || rx < 0) return ry;
            return (__ldg(prm.btab + (size_t)rx * NB + (int)(ent & 31u)) == ry) ? rx : ry;
        };
        uint32_t ent0 = ld_ent(eg), ent1 = ld_ent(eg + 2), ent2 = ld_ent(eg + 4);   // entries of my tiles T, T+2, T+4
        int ry1, rx1, ry0, rx0;
        ld_rows(ent0, ry0, rx0);
        ld_rows(ent1, ry1, rx1);                                            // rows of my next tile
        int row0 = ld_slot(ent0, ry0, rx0);                                 // slot of tile T
        const bool prof = prm.prof != nullptr;
        long long pc[4] = {0, 0, 0, 0};
        long long tp = prof ? clock64() : 0;
        for (int T = eg; T < ntiles; T += 2) {
            const int cnt = min(PB_TILE, G - PB_TILE * T);
            const int buf = T & 1;
            const uint32_t ent = ent0;
            const int row = row0;
            row0 = ld_slot(ent1, ry1, rx1);                                 // my next tile: rows were requested one iteration ago
            ent0 = ent1; ent1 = ent2;
            ld_rows(ent1, ry1, rx1);                                        // the one after
            ent2 = ld_ent(T + 6);
            if (prof) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
            mbar_wait(&dfull[buf], (uint32_t)((T >> 1) & 1));
            tc_fence_after();
            if (prof) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
            // four blocks (32 accumulator columns) per tcgen05.wait::ld; the first arg-max of a block's PB = 8 values is a
            // 3-level compare tree (the right operand wins only if strictly greater: the lowest index survives ties)
            auto emit = [&](int s, float bu, int bi) {
                const uint32_t e_s = __shfl_sync(0xffffffffu, ent, s);
                const int r_s = __shfl_sync(0xffffffffu, row, s);
                const int q = (int)(e_s & 31u);
                const unsigned long long key = ((unsigned long long)f32_ordered(bu) << 32) |
                                               (unsigned long long)(0xFFFFFFFFu - (unsigned)(PB * q + bi));
                if (jok && !(prm.dbg & 1)) __stcg(prm.bkey + (size_t)r_s * row_keys + koff + (size_t)q * J2, key);
            };
#pragma unroll
            for (int s = 0; s < PB_TILE; s += 4) {
                if (s < cnt) {                                              // warp-uniform
                    uint32_t ra[16], rb[16];
                    const bool second = s + 2 < cnt;
                    tmem_ld16(lane_addr + buf * 128 + PB * s, ra);          // blocks s, s + 1
                    if (second) tmem_ld16(lane_addr + buf * 128 + PB * (s + 2), rb);   // blocks s + 2, s + 3
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    float bu[4] = {0.f, 0.f, 0.f, 0.f};
                    int bi[4] = {0, 0, 0, 0};
                    argmax8(ra, 0, bu[0], bi[0]);
                    argmax8(ra, 8, bu[1], bi[1]);
                    if (second) { argmax8(rb, 0, bu[2], bi[2]); argmax8(rb, 8, bu[3], bi[3]); }
                    emit(s, bu[0], bi[0]);
                    if (s + 1 < cnt) emit(s + 1, bu[1], bi[1]);
                    if (s + 2 < cnt) emit(s + 2, bu[2], bi[2]);
                    if (s + 3 < cnt) emit(s + 3, bu[3], bi[3]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(&dempty[buf]); else mbar_arrive_cluster_relaxed(&dempty[buf], 0);
            }
            if (prof) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
        }
        if (prof && threadIdx.x == 0) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; o[3] = pc[3]; }
    } else if (warp == WARP_MMA2) {
        if (rank == 0) {
            // ===== MMA ISSUER (leader CTA) =====
            const uint32_t ring_addr = smem_u32(ring);
            int slot = 0;
            uint32_t sphase = 0;
            const int last_ksteps = (prm.kpad - (NCH - 1) * KCH) / 16;
            const uint32_t a_lo_off = (uint32_t)(prm.kpad / 2);
            const bool prof = prm.prof != nullptr;
            long long pc[4] = {0, 0, 0, 0};
            long long tp = prof ? clock64() : 0;
            const long long tstart = tp;
            for (int it = 0; it < ntiles; ++it) {
                const int cnt = min(PB_TILE, G - PB_TILE * it);
                const uint32_t idesc = make_idesc(256, (PB * cnt + 15) & ~15);
                const int buf = it & 1;
                if (it >= 2) mbar_wait_cluster_relaxed(&dempty[buf], (uint32_t)(((it >> 1) + 1) & 1));
                if (prof) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + D_COL0 + buf * 128;
#pragma unroll 1
                for (int kc = 0; kc < NCH; ++kc) {
                    mbar_wait(&fullL[slot], sphase);
                    if (prof) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                    mbar_wait_cluster_relaxed(&fullR[slot], sphase);
                    if (prof) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
                    tc_fence_after();
                    const uint64_t dhi = make_b_desc(ring_addr + slot * SLOT2_BYTES);
                    const uint64_t dlo = make_b_desc(ring_addr + slot * SLOT2_BYTES + MAT2_BYTES);
                    const int ksteps = (kc == NCH - 1) ? last_ksteps : KCH / 16;
                    const uint32_t a_hi0 = tmem_base + kc * (KCH / 2);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < KCH / 16; ++ks) {
                            if (ks < ksteps) {
                                const uint32_t a_hi = a_hi0 + ks * 8;
                                mma_ts2(d_tmem, a_hi, dhi + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
                                mma_ts2(d_tmem, a_hi, dlo + (uint64_t)(ks * 2), idesc, 1u);
                                mma_ts2(d_tmem, a_hi + a_lo_off, dhi + (uint64_t)(ks * 2), idesc, 1u);
                            }
                        }
                        tc_commit2(&empty[slot]);
                        if (kc == NCH - 1) tc_commit2(&dfull[buf]);
                    }
                    __syncwarp();
                    if (prof) { const long long t1 = clock64(); pc[3] += t1 - tp; tp = t1; }
                    if (++slot == NSLOT2) { slot = 0; sphase ^= 1; }
                }
            }
            if (prof && lane == 0) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[4] = pc[0]; o[5] = pc[1]; o[6] = pc[2]; o[7] = pc[3]; o[8] = clock64() - tstart; o[9] = ntiles; o[10] = b_hi - b_lo; }
        } else if (lane == 0) {
            // ===== FORWARDER (rank 1): one remote arrive per chunk =====
            int slot = 0;
            uint32_t sphase = 0;
            for (int c = 0; c < ntiles * NCH; ++c) {
                mbar_wait(&fullL[slot], sphase);
                mbar_arrive_cluster_relaxed(&fullR[slot], 0);
                if (++slot == NSLOT2) { slot = 0; sphase ^= 1; }
            }
        }
    } else if (warp < WARP_MMA2) {
        // ===== PRODUCERS (both CTAs): one row per thread; a tile = PB_TILE consecutive dirty blocks of the CTA's sequence (N rounded
        // up to a multiple of 16 rows: an odd block count leaves one pad block, produced by nobody and read by nobody), this CTA's half =====
        const int pw = warp - 4;
        const int g = lane & 7, q = lane >> 3;
        const int r = 16 * (pw >> 2) + (pw & 3) + 4 * q;      // local row 0..63
        const uint32_t t0addr = smem_u32(sT0) + 16 * g;
        const uint32_t o0 = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((((g >> 1)) ^ (r & 7)) << 4) + ((g & 1) << 3));
        const uint32_t st0 = smem_u32(ring) + o0, st1 = smem_u32(ring) + (o0 ^ 64u);
        int slot = 0;
        uint32_t phase = 0;
        // my row of tile T: block entry (chain << 5 | block) from the compact list two tiles ahead, residues one tile ahead
        auto my_row = [&](int T, bool& act, int& gr) {
            const int cntT = min(PB_TILE, G - PB_TILE * T);
            const int half = ((PB * cntT + 15) & ~15) >> 1;          // rows of this CTA (the MMA's N is a multiple of 16)
            gr = (int)rank * half + r;
            act = r < half && (gr >> PB_SHIFT) < cntT;
            if (!act) gr = 0;
        };
        auto ld_entry = [&](int T) -> uint32_t {
            if (T >= ntiles) return 0u;
            bool act; int gr;
            my_row(T, act, gr);
            return __ldg(bl + PB_TILE * T + (gr >> PB_SHIFT));
        };
        uint32_t an[5] = {0u, 0u, 0u, 0u, 0u};
        bool nact = false;
        int npos = 0, nb = 0;
        auto prep = [&](int T, uint32_t ent) {   // decode my row of tile T and put the loads of its 5 residues in flight
            nact = false;
            if (T >= ntiles) return;
            int gr;
            my_row(T, nact, gr);
            nb = (int)(ent >> 5);
            npos = PB * (int)(ent & 31u) + (gr & (PB - 1));
            const uint8_t* ap = prm.aa + (size_t)nb * prm.aa_stride + min(npos, P - 1);   // rows past the end replicate P-1
            if (!(prm.dbg & 4)) {
#pragma unroll
                for (int t = 0; t < 5; ++t) an[t] = ap[t];
            }
        };
        prep(0, ld_entry(0));
        uint32_t ent_next = ld_entry(1);
        int emit_ctr = 0;
        const bool prof = prm.prof != nullptr;
        long long pce = 0;
        const long long tstart = prof ? clock64() : 0;
        for (int T = 0; T < ntiles; ++T) {
            const bool active = nact;
            const int pos = npos;
            const int bcur = nb;
            uint32_t ra[5];
#pragma unroll
            for (int t = 0; t < 5; ++t) ra[t] = t0addr + (uint32_t)((t * PPDE_Q + (int)an[t]) * (KS * 4));
            prep(T + 1, ent_next);                             // next tile's residues in flight during this tile's chunks
            ent_next = ld_entry(T + 2);
            const bool emit_mask = (prm.r1mask != nullptr) && (emit_ctr == mp);
            if (++emit_ctr == MP) emit_ctr = 0;
            uint32_t mbits = 0u;
#pragma unroll
            for (int kc = 0; kc < NCH; ++kc) {
                const long long tw = prof ? clock64() : 0;
                mbar_wait(&empty[slot], phase ^ 1);
                if (prof) pce += clock64() - tw;
                if (active && !(prm.dbg & 2)) {
                    const uint32_t so = (uint32_t)(slot * SLOT2_BYTES);
                    if (kc == 0) produce_chunk<0>(ra, st0 + so, st1 + so, emit_mask, mbits);
                    if (kc == 1) produce_chunk<1>(ra, st0 + so, st1 + so, emit_mask, mbits);
                    if (kc == 2) produce_chunk<2>(ra, st0 + so, st1 + so, emit_mask, mbits);
                    if (kc == 3) produce_chunk<3>(ra, st0 + so, st1 + so, emit_mask, mbits);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&fullL[slot]);
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
                if (active && pos < P) {
                    int mrow = prm.rows_y ? __ldg(prm.rows_y + bcur) : prm.row_base_y + bcur;
                    if (prm.rows_x) {                          // same slot rule as the block keys
                        const int rx = __ldg(prm.rows_x + bcur);
                        if (__ldg(prm.btab + (size_t)rx * NB + (pos >> PB_SHIFT)) == mrow) mrow = rx;
                    }
                    reinterpret_cast<uint32_t*>(prm.r1mask + (((size_t)mrow * prm.m.n_nets + k) * P + pos) * 32)[g] = word;
                }
            }
        }
        if (prof && lane == 0 && pw == 0) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[12] = pce; o[13] = clock64() - tstart; }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == WARP_MMA2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// Dirty blocks of every proposal: bit q of dmask[b] is set when some conv row of block q reads a residue where y differs from x.
__global__ void __launch_bounds__(128) cnn_dirty_kernel(int n, int L, int P, int aa_stride, const uint8_t* __restrict__ aa_x,
                                                        const uint8_t* __restrict__ aa_y, uint32_t* __restrict__ dmask) {
    const int b = blockIdx.x;
    if (b >= n) return;
    __shared__ uint32_t smask;
    if (threadIdx.x == 0) smask = 0u;
    __syncthreads();
    uint32_t m = 0u;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        if (aa_x[(size_t)b * aa_stride + i] != aa_y[(size_t)b * aa_stride + i]) {
            const int p_lo = max(i - 4, 0), p_hi = min(i, P - 1);
            if (p_lo <= p_hi) m |= (1u << (p_lo >> PB_SHIFT)) | (1u << (p_hi >> PB_SHIFT));
        }
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicOr(&smask, m);
    __syncthreads();
    if (threadIdx.x == 0) dmask[b] = smask;
}

// Second half of the incremental forward, one thread per (chain, net, channel).  Nothing is copied: the proposal row's
// block table points at the current row's slots for the clean blocks and at the slots cnn_forward_inc_kernel just wrote
// for the dirty ones; the chain-level winner - the 64-bit maximum over the blocks - becomes the same `mkey` the full
// forward kernel writes.  Slot rule for a dirty block q of proposal row Y built from current row X: Y's own slot, unless
// X's table points at it (X inherited that block from an earlier state that lived in Y) - then X's own slot, which neither
// row references.  Only the two private rows of a chain and read-only fixed rows ever appear in its tables.
//
// The pool `mkey_pool` keeps, per (row, net, channel), the two largest RAW block keys of the row (before bias / relu):
//     K1 = maximum over all NB blocks (the row's winner),   K2 = maximum over the blocks other than K1's, or 0 = "not known".
// Every block outside this list has a key <= min(list).  For a proposal y built from row x with dirty-block set Dm:
//     R    = the list entries whose block is clean           (still the largest keys among the CLEAN blocks, in order)
//     cand = the new keys of the dirty blocks that are >= min(old list)   (certain to outrank every unlisted clean block)
//     new list = top-2 of R u cand                            - exact for the same reason; K2 = 0 when only one is certain
// so the merge reads 16 bytes of list + the dirty blocks' keys (2.8 of NB = 30 at pas = 2) per channel instead of all NB keys,
// fully coalesced.  Only when the new list is EMPTY (every listed block dirty and every new key below the old minimum) the clean
// blocks are rescanned: 5 % of the channels in steady state (tools/merge_stats.py) - far more than independence would give
// (0.6 %), because the proposal favours the residues the winners read.  A separate exact bound of the unlisted blocks (24-byte
// entries) did not lower that rate and measured slower (4.1 vs 3.1 ms).  max over u64 is associative and the list never holds a key that
// is not a current block key of the row: the winner is bit-identical to the full scan (tested against the full kernel).
__device__ __forceinline__ void top2_insert(unsigned long long k, unsigned long long& k1, unsigned long long& k2) {
    if (k > k1) { k2 = k1; k1 = k; }
    else if (k > k2) k2 = k;
}
// Memory-level parallelism: the dirty-block set is the same for every channel of a chain, so the CTA compacts it once
// (s_dptr) and every thread works on MERGE_E channels at a time with the list loads and the first MERGE_DP dirty-key loads
// of all of them issued before the first use (two dependent global loads per channel would otherwise bound the kernel).
// Without a list (no pool, or a full evaluation) "every block is dirty" and the same loop scans all NB keys.
constexpr int MERGE_NT = 256, MERGE_E = 3, MERGE_DP = 6;
__global__ void __launch_bounds__(MERGE_NT, 4) cnn_inc_merge_kernel(const __grid_constant__ IncParams prm) {
    const int b = blockIdx.x;
    const int J2 = 2 * prm.m.C, NB = prm.NB, nets = prm.m.n_nets;
    const uint32_t all_blocks = (NB >= 32) ? 0xFFFFFFFFu : ((1u << NB) - 1u);
    const uint32_t mask = prm.dmask ? (__ldg(prm.dmask + b) & all_blocks) : all_blocks;
    const size_t row_keys = (size_t)nets * NB * J2;
    const int ry = prm.rows_y ? __ldg(prm.rows_y + b) : prm.row_base_y + b;
    const int rx = prm.rows_x ? __ldg(prm.rows_x + b) : ry;
    const bool have_old = prm.rows_x && prm.mkey_pool && !(prm.dbg & 8);   // top-2 list of the current row available
    const uint32_t first = have_old ? mask : all_blocks;                   // blocks whose keys are read in the first round
    __shared__ int s_slot[PB_MAXNB];
    __shared__ const unsigned long long* s_ptr[PB_MAXNB];       // first key of every block (through the slot table)
    __shared__ const unsigned long long* s_dptr[PB_MAXNB];      // ... of the blocks of `first`, compacted
    __shared__ int s_nd, s_nres;
    __shared__ uint16_t s_res[PPDE_MAX_NETS * 2 * 256];                   // channels (net * J2 + j) whose clean blocks have to be rescanned
    if (threadIdx.x == 0) s_nres = 0;
    if (threadIdx.x < PB_MAXNB) {
        const int q = threadIdx.x;
        int sl = ry;
        if (q < NB && prm.rows_x) {
            const int tx = __ldg(prm.btab + (size_t)rx * NB + q);
            sl = ((mask >> q) & 1u) ? ((tx == ry) ? rx : ry) : tx;
        }
        s_slot[q] = sl;
        const unsigned long long* ptr = prm.bkey + (size_t)sl * row_keys + (size_t)(q < NB ? q : 0) * J2;
        s_ptr[q] = ptr;
        if (q < NB && ((first >> q) & 1u)) s_dptr[__popc(first & ((1u << q) - 1u))] = ptr;
        if (q == 0) s_nd = __popc(first);
    }
    __syncthreads();
    const int nd = s_nd, tot = nets * J2;
    const ulonglong2* oldp = prm.mkey_pool ? reinterpret_cast<const ulonglong2*>(prm.mkey_pool) + (size_t)rx * tot : nullptr;
    ulonglong2* newp = prm.mkey_pool ? reinterpret_cast<ulonglong2*>(prm.mkey_pool) + (size_t)ry * tot : nullptr;
    unsigned long long* outp = prm.mkey + (size_t)b * tot;
    for (int e0 = threadIdx.x; e0 < tot; e0 += MERGE_E * MERGE_NT) {
        int off[MERGE_E], kk[MERGE_E], jj[MERGE_E];
        ulonglong2 Lr[MERGE_E];
        unsigned long long kr[MERGE_E][MERGE_DP];
#pragma unroll
        for (int u = 0; u < MERGE_E; ++u) {                 // issue: list of the current row
            const int e = e0 + u * MERGE_NT;
            const int k = min(e, tot - 1) / J2, j = min(e, tot - 1) - k * J2;
            kk[u] = k; jj[u] = j;
            off[u] = k * NB * J2 + j;                       // key (k, q, j) of a row sits at q * J2 + off
            Lr[u] = (have_old && e < tot) ? __ldcg(oldp + e) : make_ulonglong2(0ull, 0ull);
        }
#pragma unroll
        for (int d = 0; d < MERGE_DP; ++d)                  // issue: the first MERGE_DP keys of the first round
#pragma unroll
            for (int u = 0; u < MERGE_E; ++u)
                kr[u][d] = (d < nd && e0 + u * MERGE_NT < tot) ? __ldcg(s_dptr[d] + off[u]) : 0ull;
#pragma unroll
        for (int u = 0; u < MERGE_E; ++u) {
            const int e = e0 + u * MERGE_NT;
            if (e < tot) {
                unsigned long long k1 = 0ull, k2 = 0ull, omin = 0ull;
                if (have_old) {
                    const ulonglong2 L = Lr[u];
                    const uint32_t q1 = (0xFFFFFFFFu - (uint32_t)(L.x & 0xFFFFFFFFull)) >> PB_SHIFT;    // blocks of the listed keys
                    const uint32_t q2 = (0xFFFFFFFFu - (uint32_t)(L.y & 0xFFFFFFFFull)) >> PB_SHIFT;
                    omin = L.y ? L.y : L.x;
                    if (!((mask >> q1) & 1u)) k1 = L.x;                                            // R: listed keys of clean blocks
                    if (L.y && !((mask >> q2) & 1u)) top2_insert(L.y, k1, k2);
                }
#pragma unroll
                for (int d = 0; d < MERGE_DP; ++d)
                    if (kr[u][d] >= omin) top2_insert(kr[u][d], k1, k2);          // (0 padding is never inserted: k1, k2 >= 0)
                for (int d0 = MERGE_DP; d0 < nd; d0 += MERGE_DP) {                // more than MERGE_DP blocks: chunks of loads
                    unsigned long long kx[MERGE_DP];
#pragma unroll
                    for (int d = 0; d < MERGE_DP; ++d) kx[d] = (d0 + d < nd) ? __ldcg(s_dptr[d0 + d] + off[u]) : 0ull;
#pragma unroll
                    for (int d = 0; d < MERGE_DP; ++d) if (kx[d] >= omin) top2_insert(kx[d], k1, k2);
                }
                if (have_old && k1 == 0ull) {               // nothing certain: the clean blocks have to be rescanned - deferred
                    s_res[atomicAdd(&s_nres, 1)] = (uint16_t)e;
                    continue;
                }
                const ppde_cnn_net_t& net = prm.m.net[kk[u]];
                outp[e] = winner_from_raw(k1, 1.f / (net.w1_scale * net.r1_scale), __ldg(net.b1 + jj[u]));
                if (newp) newp[e] = make_ulonglong2(k1, k2);
            }
        }
    }
    // Rescans, cooperatively: 5 % of the channels need the exact top-2 of all NB block keys.  Inside the loop above that was a
    // divergent branch with ceil(NB / MERGE_DP) dependent rounds of loads which nearly every warp entered (96 channels per warp
    // and iteration: 1 - 0.95^32 = 80 % per unrolled entry), i.e. ~24 dependent global-load rounds per warp - the kernel's whole
    // critical path (2.9 ms).  Now the channels are queued and 8 lanes share one: 4 independent loads per lane, three shuffle
    // levels of top-2 merges.  Order of the queue does not matter (one result per channel); max over u64 keys is exact.
    if (have_old) {
        __syncthreads();
        const int nres = s_nres;
        const int g8 = threadIdx.x & 7;
        for (int r0 = (threadIdx.x >> 5) * 4; r0 < nres; r0 += MERGE_NT / 8) {        // warp-uniform trip count (shuffles below)
            const int r = r0 + ((threadIdx.x & 31) >> 3);
            const bool live = r < nres;
            const int e = live ? s_res[r] : 0;
            const int k = e / J2, j = e - k * J2;
            const int off = k * NB * J2 + j;
            unsigned long long kx[4];
#pragma unroll
            for (int d = 0; d < 4; ++d) kx[d] = (live && g8 + 8 * d < NB) ? __ldcg(s_ptr[g8 + 8 * d] + off) : 0ull;
            unsigned long long k1 = 0ull, k2 = 0ull;
#pragma unroll
            for (int d = 0; d < 4; ++d) top2_insert(kx[d], k1, k2);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {               // merge the top-2 lists of lane pairs (block keys are distinct)
                const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
                top2_insert(o1, k1, k2);
                top2_insert(o2, k1, k2);
            }
            if (g8 == 0 && live) {
                const ppde_cnn_net_t& net = prm.m.net[k];
                outp[e] = winner_from_raw(k1, 1.f / (net.w1_scale * net.r1_scale), __ldg(net.b1 + j));
                if (newp) newp[e] = make_ulonglong2(k1, k2);
            }
        }
    }
    if (threadIdx.x < NB) prm.btab[(size_t)ry * NB + threadIdx.x] = s_slot[threadIdx.x];
}

// Dirty-block bookkeeping of the incremental forward: for chain range r = [n r / R, n (r+1) / R) (R = clusters per
// (net, channel-tile pair), the forward kernel's partition) an exclusive prefix of the chains' dirty-block counts, the
// range total, and the compact list of (chain << 4 | block) entries in sequence order.
__global__ void __launch_bounds__(1024) cnn_inc_scan_kernel(int n, int R, int NB, const uint32_t* __restrict__ dmask,
                                                            int32_t* __restrict__ boff, int32_t* __restrict__ gtot,
                                                            uint32_t* __restrict__ blist) {
    const int range = blockIdx.x;
    const int b_lo = (int)((int64_t)n * range / R), b_hi = (int)((int64_t)n * (range + 1) / R);
    const uint32_t all_blocks = (NB >= 32) ? 0xFFFFFFFFu : ((1u << NB) - 1u);
    __shared__ int wsum[32];
    __shared__ int running;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    uint32_t* out = blist + (size_t)b_lo * PB_MAXNB;
    for (int base = b_lo; base < b_hi; base += 1024) {
        const int b = base + (int)threadIdx.x;
        uint32_t mask = 0u;
        if (b < b_hi) mask = dmask ? (dmask[b] & all_blocks) : all_blocks;
        const int c = __popc(mask);
        int incl = c;                                           // inclusive warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int v = wsum[lane], iv = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, iv, o);
                if (lane >= o) iv += u;
            }
            wsum[lane] = iv - v;                                // exclusive prefix of the warp sums
        }
        __syncthreads();
        const int start = running;
        int off = start + wsum[warp] + incl - c;
        if (b < b_hi) {
            boff[b] = off;
            while (mask) {
                const int q = __ffs((int)mask) - 1;
                mask &= mask - 1u;
                out[off++] = ((uint32_t)b << 5) | (uint32_t)q;
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) running = start + wsum[31] + incl;   // last thread: total of this batch
        __syncthreads();
    }
    if (threadIdx.x == 0) gtot[range] = running;
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
// Winner lists for the backward: one block per (chain, net) decodes the 2C arg-max keys, drops channels whose max is
// 0 (relu'(0) = 0), counting-sorts them by position and orders each bucket by channel (deterministic adjoint sums).
// Record layout (uint16): start[P+1] | list[J2] (channel) | row[J2] (position), padded to `rec` entries.
__global__ void __launch_bounds__(128) cnn_winner_sort_kernel(int n_nets, int C, int P, const unsigned long long* __restrict__ mkey,
                                                              uint16_t* __restrict__ wl, int rec) {
    extern __shared__ int sw[];
    const int J2 = 2 * C;
    int* sStart = sw;                 // [P+1]
    int* sFill = sStart + (P + 1);    // [P]
    int* sPst = sFill + P;            // [J2]
    int* sList = sPst + J2;           // [J2]
    const size_t bk = blockIdx.x;     // b * n_nets + k
    const unsigned long long* keys = mkey + bk * J2;
    for (int i = threadIdx.x; i <= P; i += 128) { sStart[i] = 0; if (i < P) sFill[i] = 0; }
    __syncthreads();
    for (int j = threadIdx.x; j < J2; j += 128) {
        const unsigned long long key = keys[j];
        const float mj = __uint_as_float((unsigned)(key >> 32));
        const int pst = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu));
        const bool active = (mj > 0.f) && pst >= 0 && pst < P;
        sPst[j] = active ? pst : -1;
        if (active) atomicAdd(&sStart[pst + 1], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
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
    __syncthreads();
    for (int j = threadIdx.x; j < J2; j += 128) {
        const int pst = sPst[j];
        if (pst >= 0) sList[sStart[pst] + atomicAdd(&sFill[pst], 1)] = j;
    }
    __syncthreads();
    uint16_t* out = wl + bk * rec;
    for (int pp = threadIdx.x; pp < P; pp += 128) {
        const int s0 = sStart[pp], s1 = sStart[pp + 1];
        for (int u = s0 + 1; u < s1; ++u) {
            const int v = sList[u];
            int w = u - 1;
            while (w >= s0 && sList[w] > v) { sList[w + 1] = sList[w]; --w; }
            sList[w + 1] = v;
        }
        for (int u = s0; u < s1; ++u) { out[(P + 1) + u] = (uint16_t)sList[u]; out[(P + 1) + J2 + u] = (uint16_t)pp; }
    }
    for (int i = threadIdx.x; i <= P; i += 128) out[i] = (uint16_t)sStart[i];
}

__device__ __forceinline__ void named_bar(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

constexpr int BD_NT = 48;                // columns (touched positions) per tile of the compact delta backward
constexpr int BD_NSET = 3;               // producer sets of cnn_backward_delta_kernel: set s builds the tiles with it % BD_NSET == s
constexpr int BD_NW = 6;                 // warps per producer set (3 x 6: 18 producer warps, 23 warps in all = 80 registers)
constexpr int BD_RPW = BD_NT / BD_NW;    // columns per producer warp (8: one byte per column in the 64-bit mask words)
constexpr int BD_NTHREADS = (4 + BD_NSET * BD_NW + 1) * 32;   // warps 0-3 epilogue, 4..21 producers, 22 MMA issuer
constexpr int BD_WARP_MMA = 4 + BD_NSET * BD_NW;              // highest warp id of its scheduler (22 = 2 mod 4)
constexpr int BD_MC = 64;                // touched positions whose relu-mask bytes travel inside the record (64 bytes each)
constexpr int BD_TRAIL = 16;             // uint16 words at a FIXED place (just before the mask bytes) of a compact record:
                                         // ntile | (index of the pair list) | tstart[1 .. 14]; the (orow, cfirst) pair list
                                         // itself sits at a fixed place too, the 2 rmax words before the trailer (rmax = L + 4 tmax
                                         // rows at most) - what the fused combine (pas_reverse_accept) needs is addressable from
                                         // the chain index alone: ONE load level instead of four dependent ones.  32 bytes, so that
                                         // records stay sector-aligned
// Winner records of the DELTA backward.  For chain b and net k the per-net gradient changes between the current state x
// and the proposal y only through
//   * the conv rows whose relu mask changed: p in D0 = U_{i: x_i != y_i} [i-4, i], and
//   * the channels whose arg-max position (or liveness) changed.
// Entries (channel | side << 15; side 0 = y, 1 = x), counting-sorted by position and ordered by (side, channel) inside a
// position: every winner sitting on a position of D0 (both sides), and both ends of every moved winner.  Same record layout
// as cnn_winner_sort_kernel (start[P+1] | list[<= 2*J2]).
__global__ void __launch_bounds__(128) cnn_winner_delta_kernel(const __grid_constant__ ppde_cnn_t cm, int n_nets, int C, int P, int L, int aa_stride,
                                                               const uint8_t* __restrict__ aa_x, const uint8_t* __restrict__ aa_y,
                                                               const unsigned long long* __restrict__ mkey_y,
                                                               const unsigned long long* __restrict__ mkey_pool,
                                                               const int32_t* __restrict__ rows_x,
                                                               uint16_t* __restrict__ wl, int rec, int compact,
                                                               const uint8_t* __restrict__ r1mask, const int32_t* __restrict__ btab,
                                                               int NB, const int32_t* __restrict__ rows_y) {
    extern __shared__ int sw[];
    const int J2 = 2 * C;
    int* sStart = sw;                 // [P+1]
    int* sFill = sStart + (P + 1);    // [P]
    int* sD0 = sFill + P;             // [P]   (compact mode: re-used for the compact index of a position)
    int* sPy = sD0 + P;               // [J2] position of the y-side entry or -1
    int* sPx = sPy + J2;              // [J2]
    int* sList = sPx + J2;            // [2 J2]
    const int bk = blockIdx.x, b = bk / n_nets, k = bk - b * n_nets;
    __shared__ int sBt[2][PB_MAXNB];                  // block-table rows of the proposal / current state (mask fetch at the end)
    if (r1mask && btab && threadIdx.x < 64) {
        const int side = threadIdx.x >> 5, q = threadIdx.x & 31;
        if (q < NB) sBt[side][q] = __ldg(btab + (size_t)(side ? rows_x[b] : rows_y[b]) * NB + q);
    }
    const unsigned long long* ky = mkey_y + (size_t)bk * J2;
    const unsigned long long* kx = mkey_pool + 2 * ((size_t)rows_x[b] * n_nets + k) * J2;      // {K1, K2} per channel: K1 = raw winner
    for (int i = threadIdx.x; i <= P; i += 128) { sStart[i] = 0; if (i < P) { sFill[i] = 0; sD0[i] = 0; } }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += 128) {
        if (aa_x[(size_t)b * aa_stride + i] != aa_y[(size_t)b * aa_stride + i])
            for (int p = max(i - 4, 0); p <= min(i, P - 1); ++p) sD0[p] = 1;
    }
    __syncthreads();
    auto decode = [&](unsigned long long key) -> int {
        const float mj = __uint_as_float((unsigned)(key >> 32));
        const int pst = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu));
        return ((mj > 0.f) && pst >= 0 && pst < P) ? pst : -1;
    };
    const float unscale = 1.f / (cm.net[k].w1_scale * cm.net[k].r1_scale);
    for (int j = threadIdx.x; j < J2; j += 128) {
        // the pool holds RAW winners (cnn_inc_merge_kernel): same bias / relu epilogue as for mkey
        const int py = decode(ky[j]), px = decode(winner_from_raw(kx[2 * j], unscale, __ldg(cm.net[k].b1 + j)));
        const bool moved = py != px;
        const int ey = (py >= 0 && (moved || sD0[py])) ? py : -1;
        const int ex = (px >= 0 && (moved || sD0[px])) ? px : -1;
        sPy[j] = ey; sPx[j] = ex;
        if (ey >= 0) atomicAdd(&sStart[ey + 1], 1);
        if (ex >= 0) atomicAdd(&sStart[ex + 1], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
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
    __syncthreads();
    for (int j = threadIdx.x; j < J2; j += 128) {
        const int ey = sPy[j], ex = sPx[j];
        if (ey >= 0) sList[sStart[ey] + atomicAdd(&sFill[ey], 1)] = j;
        if (ex >= 0) sList[sStart[ex] + atomicAdd(&sFill[ex], 1)] = j | 0x8000;
    }
    __syncthreads();
    uint16_t* out = wl + (size_t)bk * rec;
    for (int pp = threadIdx.x; pp < P; pp += 128) {
        const int s0 = sStart[pp], s1 = sStart[pp + 1];
        for (int u = s0 + 1; u < s1; ++u) {           // insertion sort: y side first, channels ascending (deterministic sums)
            const int v = sList[u];
            int w = u - 1;
            while (w >= s0 && sList[w] > v) { sList[w + 1] = sList[w]; --w; }
            sList[w + 1] = v;
        }
    }
    if (!compact) {                                   // start[P+1] | list
        __syncthreads();
        for (int u = threadIdx.x; u < sStart[P]; u += 128) out[(P + 1) + u] = (uint16_t)sList[u];
        for (int i = threadIdx.x; i <= P; i += 128) out[i] = (uint16_t)sStart[i];
        return;
    }
    // compact record (only the positions that received an entry):
    //     npos | pos[npos] | start[npos+1] | list[nent] | ntile | tstart[ntile+1] | (orow, cfirst)[nr]
    // The tensor-core kernel cuts the npos columns into tiles of BD_NT; column c of a tile (position p) contributes to the
    // output rows p .. p+4 of the gradient.  Per tile: orow = the distinct output rows its columns touch (ascending),
    // cfirst = the first column of the tile (tile-relative) with pos >= row - 4; tstart = prefix of the row counts.  A row
    // touched by two tiles is listed in both (the combine adds the lists in order).
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int per = (P + 31) / 32;
        int run = 0;
        for (int i = lane * per; i < min((lane + 1) * per, P); ++i) run += (sStart[i + 1] > sStart[i]);
        int incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int base = incl - run;
        for (int i = lane * per; i < min((lane + 1) * per, P); ++i) {
            const bool has = sStart[i + 1] > sStart[i];
            sD0[i] = has ? base : -1;
            base += has;
        }
        if (lane == 31) sFill[0] = incl;              // npos (sFill is free now)
    }
    __syncthreads();
    const int npos = sFill[0];
    int* sPosC = sPy;                                 // [npos] compact position list (sPy / sPx are free now: 2 J2 >= P ints)
    for (int pp = threadIdx.x; pp < P; pp += 128) {
        const int c = sD0[pp];
        if (c >= 0) { out[1 + c] = (uint16_t)pp; out[1 + npos + c] = (uint16_t)sStart[pp]; sPosC[c] = pp; }
    }
    const int nent = sStart[P];
    if (threadIdx.x == 0) { out[0] = (uint16_t)npos; out[1 + 2 * npos] = (uint16_t)nent; }
    for (int u = threadIdx.x; u < nent; u += 128) out[2 + 2 * npos + u] = (uint16_t)sList[u];
    __syncthreads();                                  // sPosC complete
    // ---- relu-mask bytes of the first BD_MC touched positions, both sides, in the LAST BD_MC * 64 bytes of the record:
    // [c][side 0 = proposal y, 1 = current x][32 bytes].  The tensor-core kernel's producers need them before their first
    // FMA; from here they are one shared-memory read away instead of two dependent global loads (block table -> mask row).
    if (r1mask) {
        const int my = rows_y[b], mx = rows_x[b];
        uint4* mout = reinterpret_cast<uint4*>(out + rec - BD_MC * 32);
        for (int it = threadIdx.x; it < min(npos, BD_MC) * 4; it += 128) {          // 4 x 16 bytes per position: y lo, y hi, x lo, x hi
            const int c = it >> 2, part = it & 3, side = part >> 1;
            const int pp = sPosC[c];
            int row = side ? mx : my;
            if (btab) row = sBt[side][pp >> PB_SHIFT];
            mout[it] = __ldg(reinterpret_cast<const uint4*>(r1mask + (((size_t)row * n_nets + k) * P + pp) * 32) + (part & 1));
        }
    }
    // ---- output-row lists of the tiles.  Columns ascend in position, so the rows [p, p+4] of column c that no earlier column of
    // the tile covers are the last min(5, p - p_prev) of them, and for exactly those rows column c is the first contributor
    // (p_prev < row - 4): cfirst = c.  One warp scans the counts of a tile's <= BD_NT columns.
    const int ntile = (npos + BD_NT - 1) / BD_NT;
    uint16_t* oo = out + 2 + 2 * npos + nent;         // ntile | tstart[ntile+1] | (orow, cfirst)[nr]
    uint16_t* pairs = out + rec - BD_MC * 32 - BD_TRAIL - 2 * (L + 4 * ((P + BD_NT - 1) / BD_NT));   // fixed place (see BD_TRAIL)
    if (threadIdx.x == 0) { oo[0] = (uint16_t)ntile; oo[1] = 0; }
    __syncthreads();                                  // sPosC complete
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int nr_run = 0;
        for (int t = 0; t < ntile; ++t) {
            const int c0 = t * BD_NT, c1 = min(c0 + BD_NT, npos);
            int cnt[2], incl = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {             // lane owns columns c0 + 2 lane, c0 + 2 lane + 1  (BD_NT <= 64)
                const int c = c0 + 2 * lane + h;
                cnt[h] = 0;
                if (c < c1) cnt[h] = (c == c0) ? 5 : min(5, sPosC[c] - sPosC[c - 1]);
                incl += cnt[h];
            }
            const int mine = incl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int base = nr_run + incl - mine;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c0 + 2 * lane + h;
                if (c < c1) {
                    const int pp = sPosC[c];
                    for (int j = 0; j < cnt[h]; ++j) {
                        pairs[2 * (base + j)] = (uint16_t)(pp + 5 - cnt[h] + j);
                        pairs[2 * (base + j) + 1] = (uint16_t)(c - c0);
                    }
                    base += cnt[h];
                }
            }
            nr_run += __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) oo[2 + t] = (uint16_t)nr_run;
        }
    }
}

// Record builder of the compact delta backward, second version (same record, bit for bit; cnn_winner_delta_kernel stays for
// the per-position layout).  The first version bucketed the entries with a counting sort over all P positions (zeroing, atomics
// and serial scans over P-sized arrays, ten block barriers: 1.42 ms per 64k chains, 5.1 k warp-instructions per (chain, net))
// although a (chain, net) has only ~70 entries.  Here: the D0 rows are a 256-bit set; every channel emits its 0..2 entries as
// keys  position << 16 | side << 15 | channel  into an unordered list (warp ballots + one shared counter); the list is
// rank-sorted (each entry counts the smaller keys: n^2 / 128 broadcast reads per thread, one barrier); positions, starts and
// the entry list fall out of one scan over the sorted keys.
__global__ void __launch_bounds__(128, 16) cnn_delta_record_kernel(const __grid_constant__ ppde_cnn_t cm, int n_nets, int C, int P, int L, int aa_stride,
                                                               const uint8_t* __restrict__ aa_x, const uint8_t* __restrict__ aa_y,
                                                               const unsigned long long* __restrict__ mkey_y,
                                                               const unsigned long long* __restrict__ mkey_pool,
                                                               const int32_t* __restrict__ rows_x,
                                                               uint16_t* __restrict__ wl, int rec,
                                                               const uint8_t* __restrict__ r1mask, const int32_t* __restrict__ btab,
                                                               int NB, const int32_t* __restrict__ rows_y) {
    extern __shared__ int sw[];
    const int J2 = 2 * C;
    uint32_t* sKey = reinterpret_cast<uint32_t*>(sw);       // [2 J2] unordered keys
    uint32_t* sSorted = sKey + 2 * J2;                       // [2 J2]
    int* sPosC = reinterpret_cast<int*>(sSorted + 2 * J2);   // [P] compact position list
    __shared__ uint32_t sD0[8];                              // bit p: the relu mask of conv row p changed (P <= 252)
    __shared__ uint32_t sTouched[8];                         // bit p: position p carries an entry (a column of the record)
    __shared__ int sPre[9];                                  // columns before word w of sTouched; sPre[8] = npos
    __shared__ int sWoff[64];                                // first rank of every (tile, producer warp) group
    __shared__ int sCount;
    __shared__ int sBt[2][PB_MAXNB];                         // block-table rows of the proposal / current state (mask fetch)
    const int bk = blockIdx.x, b = bk / n_nets, k = bk - b * n_nets;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 8) { sD0[threadIdx.x] = 0u; sTouched[threadIdx.x] = 0u; }
    if (threadIdx.x < 64) sWoff[threadIdx.x] = 0x7fffffff;
    if (threadIdx.x == 8) sCount = 0;
    if (r1mask && btab && threadIdx.x >= 64) {
        const int side = (threadIdx.x - 64) >> 5, q = threadIdx.x & 31;
        if (q < NB) sBt[side][q] = __ldg(btab + (size_t)(side ? rows_x[b] : rows_y[b]) * NB + q);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += 128) {
        if (aa_x[(size_t)b * aa_stride + i] != aa_y[(size_t)b * aa_stride + i])
            for (int pp = max(i - 4, 0); pp <= min(i, P - 1); ++pp) atomicOr(&sD0[pp >> 5], 1u << (pp & 31));
    }
    __syncthreads();
    const unsigned long long* ky = mkey_y + (size_t)bk * J2;
    const unsigned long long* kx = mkey_pool + 2 * ((size_t)rows_x[b] * n_nets + k) * J2;      // {K1, K2} per channel: K1 = raw winner
    const float unscale = 1.f / (cm.net[k].w1_scale * cm.net[k].r1_scale);
    auto decode = [&](unsigned long long key) -> int {
        const float mj = __uint_as_float((unsigned)(key >> 32));
        const int pst = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu));
        return ((mj > 0.f) && pst >= 0 && pst < P) ? pst : -1;
    };
    for (int j0 = 0; j0 < J2; j0 += 128) {
        const int j = j0 + threadIdx.x;
        int ey = -1, ex = -1;
        if (j < J2) {
            const int py = decode(ky[j]), px = decode(winner_from_raw(kx[2 * j], unscale, __ldg(cm.net[k].b1 + j)));
            const bool moved = py != px;
            if (py >= 0 && (moved || ((sD0[py >> 5] >> (py & 31)) & 1u))) ey = py;
            if (px >= 0 && (moved || ((sD0[px >> 5] >> (px & 31)) & 1u))) ex = px;
        }
        const uint32_t by = __ballot_sync(0xffffffffu, ey >= 0), bx = __ballot_sync(0xffffffffu, ex >= 0);
        int base = 0;
        if (lane == 0 && (by | bx)) base = atomicAdd(&sCount, __popc(by) + __popc(bx));
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t below = (1u << lane) - 1u;
        if (ey >= 0) { sKey[base + __popc(by & below)] = ((uint32_t)ey << 16) | (uint32_t)j; atomicOr(&sTouched[ey >> 5], 1u << (ey & 31)); }
        if (ex >= 0) { sKey[base + __popc(by) + __popc(bx & below)] = ((uint32_t)ex << 16) | 0x8000u | (uint32_t)j; atomicOr(&sTouched[ex >> 5], 1u << (ex & 31)); }
    }
    __syncthreads();
    const int nent = sCount;
    // columns = touched positions in ascending order: column of position p = number of touched positions below p
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < 8; ++w) { sPre[w] = run; run += __popc(sTouched[w]); }
        sPre[8] = run;
    }
    __syncthreads();
    const int npos = sPre[8];
    const int ntile = (npos + BD_NT - 1) / BD_NT;
    const int nw = BD_NW * ntile + 1;
    auto col_of = [&](int pp) -> int { return sPre[pp >> 5] + __popc(sTouched[pp >> 5] & ((1u << (pp & 31)) - 1u)); };
    // re-key every entry with its place in the producers' work order: tile t = c / BD_NT, producer warp w = (c % BD_NT) % BD_NW,
    // slot sl = (c % BD_NT) / BD_NW of that warp, then (side, channel) as before:   key = (48 t + 8 w + sl) << 16 | side << 15 | sl << 12 | channel
    // The rank sort then yields the list each producer warp of cnn_backward_delta_kernel walks front to back (it used to
    // flatten the entries of its 6 strided columns itself: shuffles, prefix sums and a staging list per tile and warp - 12 % of
    // that kernel's instructions, on the producers' critical path).  Within a column the order (side, channel) is unchanged.
    for (int i = threadIdx.x; i < nent; i += 128) {
        const uint32_t ki = sKey[i];
        const int c = col_of((int)(ki >> 16));
        const int t = c / BD_NT, r = c - t * BD_NT, w = r % BD_NW, sl = r / BD_NW;
        sKey[i] = ((uint32_t)(t * BD_NT + w * BD_RPW + sl) << 16) | (ki & 0x81FFu) | ((uint32_t)sl << 12);
    }
    __syncthreads();
    // record: npos | pos[npos] | woff[8 ntile + 1] | list[nent] | (fixed places) pairs | trailer | masks
    uint16_t* out = wl + (size_t)bk * rec;
    for (int i = threadIdx.x; i < nent; i += 128) {          // rank sort (keys are distinct): entry = low half of the key
        const uint32_t ki = sKey[i];
        int rank = 0;
        for (int q = 0; q < nent; ++q) rank += (sKey[q] < ki);
        out[1 + npos + nw + rank] = (uint16_t)(ki & 0xFFFFu);
        const int gq = (int)(ki >> 16);                      // group (tile, warp) = 8 t + w: its first entry has the smallest rank
        atomicMin(&sWoff[BD_NW * (gq / BD_NT) + (gq % BD_NT) / BD_RPW], rank);
    }
    __syncthreads();
    // woff[x] = entries before group x = first rank of the next non-empty group (suffix minimum; nent behind the last one)
    if (threadIdx.x < 32) {
        int a0 = (2 * lane < nw) ? sWoff[2 * lane] : nent, a1 = (2 * lane + 1 < nw) ? sWoff[2 * lane + 1] : nent;   // nw <= 8 * 7 + 1 <= 64
        a0 = min(a0, a1);
        int run = a0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_down_sync(0xffffffffu, run, o);
            if (lane + o < 32) run = min(run, v);
        }
        const int nxt = __shfl_down_sync(0xffffffffu, run, 1);                // suffix minimum of the lanes behind me
        if (2 * lane + 1 < nw) out[1 + npos + 2 * lane + 1] = (uint16_t)min(a1, lane < 31 ? nxt : nent);
        if (2 * lane < nw) out[1 + npos + 2 * lane] = (uint16_t)run;
    }
    for (int pp = threadIdx.x; pp < P; pp += 128) {
        if ((sTouched[pp >> 5] >> (pp & 31)) & 1u) { const int c = col_of(pp); sPosC[c] = pp; out[1 + c] = (uint16_t)pp; }
    }
    if (threadIdx.x == 0) out[0] = (uint16_t)npos;
    __syncthreads();
    // ---- relu-mask bytes of the first BD_MC touched positions, both sides, in the LAST BD_MC * 64 bytes of the record
    if (r1mask) {
        const int my = rows_y[b], mx = rows_x[b];
        uint4* mout = reinterpret_cast<uint4*>(out + rec - BD_MC * 32);
        for (int it = threadIdx.x; it < min(npos, BD_MC) * 4; it += 128) {          // 4 x 16 bytes per position: y lo, y hi, x lo, x hi
            const int c = it >> 2, part = it & 3, side = part >> 1;
            const int pp = sPosC[c];
            int row = side ? mx : my;
            if (btab) row = sBt[side][pp >> PB_SHIFT];
            mout[it] = __ldg(reinterpret_cast<const uint4*>(r1mask + (((size_t)row * n_nets + k) * P + pp) * 32) + (part & 1));
        }
    }
    // ---- output-row lists of the tiles (see cnn_winner_delta_kernel)
    uint16_t* pairs = out + rec - BD_MC * 32 - BD_TRAIL - 2 * (L + 4 * ((P + BD_NT - 1) / BD_NT));   // fixed place (see BD_TRAIL)
    __shared__ __align__(16) uint16_t sTr[BD_TRAIL];
    if (threadIdx.x < 32) {
        if (lane < BD_TRAIL) sTr[lane] = 0;
        __syncwarp();
        int nr_run = 0;
        for (int t = 0; t < ntile; ++t) {
            const int c0 = t * BD_NT, c1 = min(c0 + BD_NT, npos);
            int cnt[2], incl = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {             // lane owns columns c0 + 2 lane, c0 + 2 lane + 1  (BD_NT <= 64)
                const int c = c0 + 2 * lane + h;
                cnt[h] = 0;
                if (c < c1) cnt[h] = (c == c0) ? 5 : min(5, sPosC[c] - sPosC[c - 1]);
                incl += cnt[h];
            }
            const int mine = incl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            int base = nr_run + incl - mine;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = c0 + 2 * lane + h;
                if (c < c1) {
                    const int pp = sPosC[c];
                    for (int j = 0; j < cnt[h]; ++j) {
                        pairs[2 * (base + j)] = (uint16_t)(pp + 5 - cnt[h] + j);
                        pairs[2 * (base + j) + 1] = (uint16_t)(c - c0);
                    }
                    base += cnt[h];
                }
            }
            nr_run += __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0 && t < BD_TRAIL - 2) sTr[2 + t] = (uint16_t)nr_run;        // tstart[t + 1]
        }
        if (lane == 0) { sTr[0] = (uint16_t)ntile; sTr[1] = (uint16_t)nent; }
        __syncwarp();
        // the fixed trailer, one full 32-byte sector (two 16-byte stores)
        if (lane < 2) reinterpret_cast<uint4*>(out + rec - BD_MC * 32 - BD_TRAIL)[lane] = reinterpret_cast<const uint4*>(sTr)[lane];
    }
}

// Delta backward, second kernel: the proposal's gradient row from the current state's row, the change of the Potts field and the
// SPARSE per-net changes the tensor-core kernel left in the scratch (rows listed in the records, values [row][20]):
//     G_y = G_x + (Gp_y - Gp_x)(window);   then for k = 0 .. n_nets-1, tile by tile, row by row:  G_y[row] += lamda/n_nets * dGc_k[row]
// One CTA per chain, the row is assembled in shared memory; fixed order (net, tile, row): deterministic.
__global__ void __launch_bounds__(256) cnn_grad_combine_sparse_kernel(int n, int L, int n_nets, float scale, ppde_potts_t pm,
                                                                      const float* __restrict__ vals, int vcap,
                                                                      const uint16_t* __restrict__ wl, int rec,
                                                                      const float* __restrict__ Gp, int64_t Gp_stride,
                                                                      float* __restrict__ G, int64_t G_stride,
                                                                      const int32_t* __restrict__ rows_x, const int32_t* __restrict__ rows_y) {
    extern __shared__ __align__(16) float srow[];     // [20 L]
    const int b = blockIdx.x, NE = L * PPDE_Q;
    const int wlo = pm.win_lo * PPDE_Q, whi = (pm.win_lo + pm.Lp) * PPDE_Q;     // multiples of 4
    const int rx = rows_x[b], ry = rows_y[b];
    const float4* gx = reinterpret_cast<const float4*>(G + (int64_t)rx * G_stride);
    const float* px = Gp ? Gp + (int64_t)rx * Gp_stride : nullptr;
    const float* py = Gp ? Gp + (int64_t)ry * Gp_stride : nullptr;
    float4* s4 = reinterpret_cast<float4*>(srow);
    for (int q = threadIdx.x; q < NE / 4; q += blockDim.x) {
        float4 base = __ldcs(gx + q);
        const int e = q * 4;
        if (Gp && e >= wlo && e < whi) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(py + (e - wlo)));
            const float4 c = __ldcs(reinterpret_cast<const float4*>(px + (e - wlo)));
            base.x += a.x - c.x; base.y += a.y - c.y; base.z += a.z - c.z; base.w += a.w - c.w;
        }
        s4[q] = base;
    }
    __syncthreads();
    for (int k = 0; k < n_nets; ++k) {
        const uint16_t* r = wl + ((size_t)b * n_nets + k) * rec;
        const uint16_t* tr = r + rec - BD_MC * 32 - BD_TRAIL;                              // trailer: ntile | nent | tstart[1 ..]
        const int ntile = tr[0];
        const uint16_t* pairs = r + rec - BD_MC * 32 - BD_TRAIL - 2 * (vcap / PPDE_Q);      // (orow, cfirst) per output row, fixed place
        const float* v = vals + ((size_t)k * n + b) * vcap;
        for (int t = 0; t < ntile; ++t) {
            const int r0 = t ? tr[1 + t] : 0, r1 = tr[2 + t];
            for (int it = r0 * PPDE_Q + (int)threadIdx.x; it < r1 * PPDE_Q; it += blockDim.x) {
                const int rr = it / PPDE_Q, a = it - rr * PPDE_Q;
                const int e = (int)pairs[2 * rr] * PPDE_Q + a;
                srow[e] = fmaf(scale, __ldcs(v + it), srow[e]);
            }
            __syncthreads();
        }
    }
    float4* gy = reinterpret_cast<float4*>(G + (int64_t)ry * G_stride);
    for (int q = threadIdx.x; q < NE / 4; q += blockDim.x) gy[q] = s4[q];
}

// G_y = G_x + (Gp_y - Gp_x)(window) + lamda / n_nets * (dGc_0 + dGc_1 + dGc_2)   (delta backward; fixed summation order)
__global__ void cnn_grad_combine_delta_kernel(int n, int NE, int n_nets, float scale, ppde_potts_t pm,
                                              const float* __restrict__ Gc, const float* __restrict__ Gp, int64_t Gp_stride,
                                              float* __restrict__ G, int64_t G_stride,
                                              const int32_t* __restrict__ rows_x, const int32_t* __restrict__ rows_y) {
    const int b = blockIdx.x;
    const int wlo = pm.win_lo * PPDE_Q, whi = (pm.win_lo + pm.Lp) * PPDE_Q;     // multiples of 4
    const int rx = rows_x[b], ry = rows_y[b];
    const float4* gx = reinterpret_cast<const float4*>(G + (int64_t)rx * G_stride);
    float4* gy = reinterpret_cast<float4*>(G + (int64_t)ry * G_stride);
    const float* px = Gp ? Gp + (int64_t)rx * Gp_stride : nullptr;
    const float* py = Gp ? Gp + (int64_t)ry * Gp_stride : nullptr;
    for (int q = threadIdx.x; q < NE / 4; q += blockDim.x) {
        float4 acc = __ldcs(reinterpret_cast<const float4*>(Gc + (size_t)b * NE) + q);
        for (int k = 1; k < n_nets; ++k) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(Gc + ((size_t)k * n + b) * NE) + q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float4 base = gx[q];
        const int e = q * 4;
        if (Gp && e >= wlo && e < whi) {
            const float4 a = *reinterpret_cast<const float4*>(py + (e - wlo));
            const float4 c = *reinterpret_cast<const float4*>(px + (e - wlo));
            base.x += a.x - c.x; base.y += a.y - c.y; base.z += a.z - c.z; base.w += a.w - c.w;
        }
        gy[q] = make_float4(fmaf(scale, acc.x, base.x), fmaf(scale, acc.y, base.y), fmaf(scale, acc.z, base.z),
                            fmaf(scale, acc.w, base.w));
    }
}

constexpr int BW_NT = 64;               // positions per tile (N of the MMA)
constexpr int BW_NT_PROD = 512;         // 16 producer warps in two sets of 8: set s builds the tiles with it % 2 == s, 8 rows per warp
constexpr int BW_NTHREADS = NT_EPI + BW_NT_PROD + 32;   // 672: warps 0-3 epilogue, 4-19 producers, 20 MMA issuer
constexpr int BW_WARP_MMA = 20;         // highest warp id of its scheduler: top arbitration priority
constexpr int BW_MAT = BW_NT * KCH * 2; // one [64 x 64] fp16 operand matrix (8 KB)
constexpr int BW_SLOT = 2 * BW_MAT;     // hi + lo
constexpr int BW_MAXCH = 4;             // K chunks per tile (kpad <= 256)
constexpr int BW_NDBUF = 4;             // accumulator buffers of 64 TMEM columns
constexpr int BW_NBUF = 3;              // operand tile buffers in shared memory (3 x 64 KB): a producer set starts storing its next
                                        // tile while the MMA still reads the set's previous one

struct BwdParams {
    ppde_cnn_t m;
    ppde_potts_t pm;
    const uint8_t* aa;
    int aa_stride;
    int n;
    const unsigned long long* mkey;
    const uint8_t* r1mask;              // [rows, n_nets, P, 32] relu mask bits written by the forward kernel
    const int32_t* mask_rows;           // [n] row of chain b in r1mask, NULL = mask_row_base + b
    int mask_row_base;
    const int32_t* mask_rows_x;         // delta mode only: [n] r1mask row of the chain's CURRENT state
    int nrec;                           // compact delta kernel: record buffers in shared memory
    const int32_t* btab; int NB;        // optional block table of the pools [rows, NB]: the mask rows of block q of row r live in row btab[r][q]
    const uint16_t* wl; int rec;        // winner records from cnn_winner_sort_kernel
    float* Gc;                          // [n_nets][n][20L] per-net partial gradients (combined by cnn_grad_combine_kernel);
                                        // compact delta kernel: [n_nets][n][vcap] sparse values [row][20] of the rows the record lists
    int vcap;                           // floats per (net, chain) of the sparse scratch
    int ctas_per_net;
    int tiles_per_chain, nch, kpad;
    int dbg;                            // profiling experiments only (PPDE_BWD_DEBUG): 2 = skip gathers
    long long* prof;                    // optional [grid][16] cycle counters (instrumented build)
};

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) { tmem_ld32(taddr, r); }
// explicit shared-space scalar loads (32-bit addresses; the record / decoder pointers lose their address space through
// the alignment casts and would otherwise compile to generic LD with 64-bit address math and long-scoreboard waits)
__device__ __forceinline__ int lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return (int)v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// Backward of one net for a contiguous block of chains, persistent CTA.
//
//   GEMM  Y[m(a,t), p] = sum_c W0[c,a,t] * A[p,c]      M = 128 rows (100 used), N = 64 positions per tile, K = C
//   A[p,c] = 1[r1[p,c] > 0] * sum_{j in winners(p)} d_j W1[j,c]       (adjoint of the conv output)
//
// * A-operand (W0^T, fp16 hi/lo) is resident in tensor memory.  Row order m(a,t) = 32*(a/6) + 5*(a%6) + t puts the 5
//   taps of a residue in 5 adjacent lanes of one warp, so the col2im below is a warp-shuffle reduction in registers.
// * Producers: two sets of 8 warps alternate over the tiles (set s <-> ring buffer s); a warp owns 8 rows of its set's
//   tile, lane l owns channels 8l..8l+7.  The winners of those rows are one contiguous run of the chain's (position,
//   channel)-sorted list (staged in shared memory by a bulk copy one chain ahead, read with explicit ld.shared); a warp
//   streams their W1 rows from L2 in groups of 4 (8 x 16-byte loads in flight per lane, the first group issued BEFORE the
//   tile buffer is waited for), rows without winners are stored as zeros without arithmetic, accumulates
//   d_j * W1[j,:] in registers in list order (deterministic), and on every row boundary applies the relu mask (bits
//   from the forward), the power-of-two scale and the fp16 hi/lo split and writes 16 + 16 bytes straight into the
//   K-major SW128 operand ring.  All control flow is warp-uniform.
// * Epilogue: tcgen05.ld (lane = (a,t) row, 64 columns), then  G[p0+i, a] = sum_t Y[m(a,t), i-t]  by 64 shuffles: lane
//   d of a residue's 5-lane group accumulates the outputs i = d (mod 5); 4 partial outputs carry into the next tile.
//   Results go to a shared [20L] row and are flushed per chain with coalesced 16-byte stores.
// DELTA = true: the records come from cnn_winner_delta_kernel (bit 15 of an entry = side: 0 proposal y, 1 current state x) and
// the operand rows are  dA[p,:] = mask_y[p] . sum_{y-side} d_j W1[j,:] - mask_x[p] . sum_{x-side} d_j W1[j,:]  (mask applied per
// entry, one accumulator), so the kernel produces the CHANGE of the per-net gradient between the current state and the proposal.
template <bool PROF, bool DELTA>
__global__ void __launch_bounds__(BW_NTHREADS, 1) cnn_backward_tc_kernel(const __grid_constant__ BwdParams prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int C = prm.m.C, P = prm.m.P, L = prm.m.L, J2 = 2 * C, NE = L * PPDE_Q;
    const int nch = prm.nch;
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* sGc = reinterpret_cast<float*>(ring + BW_NBUF * BW_MAXCH * BW_SLOT);   // [NE] chain gradient row
    float* sDj = sGc + ((NE + 3) & ~3);                                      // [J2] decoder weights
    uint16_t* sRec = reinterpret_cast<uint16_t*>((reinterpret_cast<uintptr_t>(sDj + J2) + 15) & ~(uintptr_t)15);   // [2][rec]
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sRec + 2 * prm.rec) + 7) & ~(uintptr_t)7);
    uint64_t* full = bars;                      // [BW_NBUF] tile buffers: producers -> MMA (8 warp arrivals: one producer set)
    uint64_t* empty = full + BW_NBUF;           // [BW_NBUF] MMA -> producers
    uint64_t* dfull = empty + BW_NBUF;          // [BW_NDBUF] MMA -> epilogue
    uint64_t* dempty = dfull + BW_NDBUF;        // [BW_NDBUF] epilogue -> MMA (4 warp arrivals)
    uint64_t* recfull = dempty + BW_NDBUF;      // [2] winner record landed (bulk copy, tx bytes)
    uint64_t* recempty = recfull + 2;           // [2] producers done with the record (16 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(recempty + 2);

    const int k = blockIdx.x / prm.ctas_per_net;                       // this CTA's net
    const int within = blockIdx.x - k * prm.ctas_per_net;
    if (k >= prm.m.n_nets) return;
    const int b_lo = (int)((int64_t)prm.n * within / prm.ctas_per_net);
    const int b_hi = (int)((int64_t)prm.n * (within + 1) / prm.ctas_per_net);
    const int tpc = prm.tiles_per_chain;
    const int nchains = b_hi - b_lo;
    const int ntiles = nchains * tpc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ppde_cnn_net_t net = prm.m.net[k];
    const uint32_t rec_bytes = (uint32_t)prm.rec * 2u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < BW_NBUF; ++s) { mbar_init(&full[s], BW_NT_PROD / 64); mbar_init(&empty[s], 1); }   // 8 warps of one set per tile
        for (int s = 0; s < 2; ++s) { mbar_init(&recfull[s], 1); mbar_init(&recempty[s], BW_NT_PROD / 32); }
        for (int d = 0; d < BW_NDBUF; ++d) { mbar_init(&dfull[d], 1); mbar_init(&dempty[d], NT_EPI / 32); }
        fence_barrier_init();
    }
    if (warp == BW_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int j = threadIdx.x; j < J2; j += BW_NTHREADS) sDj[j] = net.d[j];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {   // A = W0^T (scaled, fp16 hi/lo) -> TMEM: lane m(a,t) = 32*(a/6) + 5*(a%6) + t
        const int grp = lane / 5, t = lane - 5 * grp, a = 6 * warp + grp;
        const bool rowok = lane < 30 && a < PPDE_Q;
        const int nrow = t * PPDE_Q + a;                                // W0r[c][t][a]
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int ks = 0; ks < prm.kpad / 16; ++ks) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c0 = ks * 16 + 2 * q;
                const float w0 = (rowok && c0 < C) ? net.W0r[(size_t)c0 * 100 + nrow] * net.w0_scale : 0.f;
                const float w1 = (rowok && c0 + 1 < C) ? net.W0r[(size_t)(c0 + 1) * 100 + nrow] * net.w0_scale : 0.f;
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
        // ===== EPILOGUE: lane = (a, t); col2im by shuffles inside the residue's 5-lane group =====
        const int grp = lane / 5, d = lane - 5 * grp, a = 6 * warp + grp;
        const bool rowok = lane < 30 && a < PPDE_Q;
        const int gbase = 5 * grp;
        const int tid = threadIdx.x;
        const float unscale = 1.f / (net.w0_scale * net.adj_scale);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL0;
        int srcv[5];                      // source lane of column c = 5u+v: tap t' = (d - v) mod 5 of my group
#pragma unroll
        for (int v = 0; v < 5; ++v) srcv[v] = min(gbase + (d - v + 5) % 5, 31);
        float carry = 0.f;                // partial output i = d (< 4) of the NEXT tile (from this tile's columns 60..63)
        long long pc[4] = {0, 0, 0, 0};
        long long tp = PROF ? clock64() : 0;
        int b = b_lo, tn = 0;
        for (int it = 0; it < ntiles; ++it) {
            const int buf = it & (BW_NDBUF - 1);
            const int p0 = tn * BW_NT;
            mbar_wait(&dfull[buf], (uint32_t)((it / BW_NDBUF) & 1));
            if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
            tc_fence_after();
            uint32_t y[64];
            tmem_ld32(lane_addr + buf * BW_NT, y);
            tmem_ld32(lane_addr + buf * BW_NT + 32, y + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dempty[buf]);
            if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
            // acc[s] = output i = 5 s + d of this tile (relative position), contributions in order of increasing column
            float acc[14];
#pragma unroll
            for (int s = 0; s < 14; ++s) acc[s] = 0.f;
            acc[0] = (tn == 0) ? 0.f : carry;
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                const int u = c / 5, v = c - 5 * u;
                const float x = __shfl_sync(0xffffffffu, __uint_as_float(y[c]), srcv[v]);
                if (d >= v) acc[u] += x; else acc[u + 1] += x;
            }
            // outputs i < 64 are complete; i = 64..67 (lanes d = 4 slot 12; d = 0,1,2 slot 13) carry into the next tile,
            // where output i - 64 lives in lane d' = i - 64:  d' = 0 <- (lane 4, slot 12),  d' = 1,2,3 <- (lane d'-1, slot 13)
            {
                const float c12 = __shfl_sync(0xffffffffu, acc[12], min(gbase + 4, 31));
                const float c13 = __shfl_sync(0xffffffffu, acc[13], max(gbase + d - 1, 0));
                carry = (d == 0) ? c12 : ((d < 4) ? c13 : 0.f);
            }
            if (rowok) {
#pragma unroll
                for (int s = 0; s < 13; ++s) {
                    const int i = 5 * s + d;
                    if (i < 64 && p0 + i < L) sGc[(p0 + i) * PPDE_Q + a] = acc[s] * unscale;
                }
                if (tn == tpc - 1 && d < 4 && p0 + 64 + d < L) sGc[(p0 + 64 + d) * PPDE_Q + a] = carry * unscale;
            }
            if (PROF) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
            if (tn == tpc - 1) {
                // flush the chain's partial gradient (streaming 16-byte stores)
                named_bar(2, NT_EPI);
                float4* dst = reinterpret_cast<float4*>(prm.Gc + ((size_t)k * prm.n + b) * NE);
                const float4* src = reinterpret_cast<const float4*>(sGc);
                for (int e = tid; e < NE / 4; e += NT_EPI) __stcs(dst + e, src[e]);
                named_bar(2, NT_EPI);
                if (PROF) { const long long t1 = clock64(); pc[3] += t1 - tp; tp = t1; }
                tn = 0; ++b;
            } else {
                ++tn;
            }
        }
        if (PROF && prm.prof && threadIdx.x == 0) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; o[3] = pc[3]; }
    } else if (warp == BW_WARP_MMA) {
        // ===== MMA ISSUER: warp-uniform loop, one elected lane issues; also stages the winner records (bulk copies) =====
        const uint32_t idesc = make_idesc(128, BW_NT);
        const uint32_t ring_addr = smem_u32(ring);
        const int last_ksteps = (prm.kpad - (nch - 1) * KCH) / 16;
        const uint32_t a_lo_off = (uint32_t)(prm.kpad / 2);
        long long pc[3] = {0, 0, 0};
        long long tp = PROF ? clock64() : 0;
        auto stage_record = [&](int ci) {      // chain ci of this CTA -> record buffer ci & 1
            const int rb = ci & 1;
            if (ci >= 2) mbar_wait(&recempty[rb], (uint32_t)(((ci >> 1) + 1) & 1));
            if (elect_one()) {
                mbar_expect_tx(&recfull[rb], rec_bytes);
                bulk_g2s(sRec + (size_t)rb * prm.rec, prm.wl + ((size_t)(b_lo + ci) * prm.m.n_nets + k) * prm.rec, rec_bytes,
                         &recfull[rb]);
            }
            __syncwarp();
        };
        if (nchains > 0) stage_record(0);
        int tn = 0, ci = 0;
        for (int it = 0; it < ntiles; ++it) {
            const int tb = it % BW_NBUF;                             // tile ring buffer
            const int buf = it & (BW_NDBUF - 1);                      // accumulator buffer
            if (tn == 0 && ci + 1 < nchains) stage_record(ci + 1);   // one chain ahead
            if (it >= BW_NDBUF) mbar_wait(&dempty[buf], (uint32_t)(((it / BW_NDBUF) + 1) & 1));
            if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
            mbar_wait(&full[tb], (uint32_t)((it / BW_NBUF) & 1));
            if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + D_COL0 + buf * BW_NT;
            if (elect_one()) {
                for (int kc = 0; kc < nch; ++kc) {
                    const int slot = tb * BW_MAXCH + kc;
                    const uint64_t dhi = make_b_desc(ring_addr + slot * BW_SLOT);
                    const uint64_t dlo = make_b_desc(ring_addr + slot * BW_SLOT + BW_MAT);
                    const int ksteps = (kc == nch - 1) ? last_ksteps : KCH / 16;
                    const uint32_t a_hi0 = tmem_base + kc * (KCH / 2);
#pragma unroll
                    for (int ks = 0; ks < KCH / 16; ++ks) {
                        if (ks < ksteps) {
                            const uint32_t a_hi = a_hi0 + ks * 8;
                            mma_ts(d_tmem, a_hi, dhi + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
                            mma_ts(d_tmem, a_hi, dlo + (uint64_t)(ks * 2), idesc, 1u);
                            mma_ts(d_tmem, a_hi + a_lo_off, dhi + (uint64_t)(ks * 2), idesc, 1u);
                        }
                    }
                }
                tc_commit(&empty[tb]);
                tc_commit(&dfull[buf]);
            }
            __syncwarp();
            if (PROF) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
            if (++tn == tpc) { tn = 0; ++ci; }
        }
        if (PROF && prm.prof && lane == 0) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[4] = pc[0]; o[5] = pc[1]; o[6] = pc[2]; }
    } else {
        // ===== PRODUCERS =====
        // Two sets of 8 warps; set s = pw / 8 builds the tiles with it % 2 == s (always in ring buffer s), warp pw % 8 owns rows
        // 8 (pw % 8) .. + 7 of them.  A warp's work on one tile is a latency chain (record reads -> W1 rows from L2 -> row
        // stores); with each set working on its own tile two such chains overlap per SM.
        const int pw = warp - 4;
        const int pset = pw >> 3;
        const int r0 = 8 * (pw & 7);
        const bool lact = 8 * lane < prm.kpad;
        const float adj_scale = net.adj_scale;
        const float* wbase = net.W1p + 8 * lane;
        // element (row r, k = 8 lane + e): chunk lane/8, 16-byte unit (lane%8) ^ (r%8) of the row's 128 bytes
        const uint32_t ring_lane = smem_u32(ring) + (uint32_t)((lane >> 3) * BW_SLOT);
        const uint32_t unit = (uint32_t)(lane & 7);
        const uint32_t rec_a = smem_u32(sRec), dj_a = smem_u32(sDj);
        long long pc[4] = {0, 0, 0, 0};
        long long tp = PROF ? clock64() : 0;
        // relu-mask bytes of my 8 rows of a tile (byte rr of the 64-bit word = row r0 + rr): proposal rows and, in delta mode,
        // the current state's rows
        auto load_masks = [&](int bb, int pp0, unsigned long long& my, unsigned long long& mx) {
            my = 0ull; mx = 0ull;
            int mr = prm.mask_rows ? __ldg(prm.mask_rows + bb) : prm.mask_row_base + bb;
            const int blk = min((pp0 + r0) >> PB_SHIFT, prm.NB - 1);     // my 8 rows (aligned to 8) lie in one block of PB positions
            if (prm.btab) mr = __ldg(prm.btab + (size_t)mr * prm.NB + blk);
            const uint8_t* mrow = prm.r1mask + (((size_t)mr * prm.m.n_nets + k) * P + pp0 + r0) * 32 + lane;
#pragma unroll
            for (int rr = 0; rr < 8; ++rr)
                if (lact && pp0 + r0 + rr < P) my |= (unsigned long long)__ldg(mrow + rr * 32) << (8 * rr);
            if (DELTA) {
                int mxr = __ldg(prm.mask_rows_x + bb);
                if (prm.btab) mxr = __ldg(prm.btab + (size_t)mxr * prm.NB + blk);
                const uint8_t* xrow = prm.r1mask + (((size_t)mxr * prm.m.n_nets + k) * P + pp0 + r0) * 32 + lane;
#pragma unroll
                for (int rr = 0; rr < 8; ++rr)
                    if (lact && pp0 + r0 + rr < P) mx |= (unsigned long long)__ldg(xrow + rr * 32) << (8 * rr);
            }
        };
        int tn = 0, ci = 0;
        for (int it = 0; it < ntiles; ++it) {
            const int rb = ci & 1;
            if ((it & 1) == pset) {
                const int tb = it % BW_NBUF;
                const int p0 = tn * BW_NT;
                unsigned long long m8 = 0ull, m8x = 0ull;
                load_masks(b_lo + ci, p0, m8, m8x);                    // independent of the winners: in flight during the record reads
                mbar_wait(&recfull[rb], (uint32_t)((ci >> 1) & 1));
                const uint32_t rs = rec_a + (uint32_t)rb * rec_bytes;         // this chain's record: start[P+1] | list
                const uint32_t ls = rs + 2u * (uint32_t)(P + 1);
                const uint32_t tile_addr = ring_lane + (uint32_t)(tb * BW_MAXCH * BW_SLOT);
                int e = lds_u16(rs + 2u * (uint32_t)min(p0 + r0, P));
                const int eB = (prm.dbg & 2) ? e : lds_u16(rs + 2u * (uint32_t)min(p0 + r0 + 8, P));
                int cur = 0;
                int rend = lds_u16(rs + 2u * (uint32_t)min(p0 + r0 + 1, P));
                bool have_buf = false;      // the tile buffer is waited for at the first row store: the W1 loads of the first
                                            // group are already in flight by then
                bool dirty_row = false;     // the current row received an entry
                float acc[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                auto store_row = [&]() {   // row r0 + cur <- mask * scale * acc, fp16 hi (truncated: exact) + lo
                    if (!have_buf) {
                        if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                        mbar_wait(&empty[tb], (uint32_t)(((it / BW_NBUF) + 1) & 1));
                        have_buf = true;
                        if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
                    }
                    const int r = r0 + cur;
                    const uint32_t addr = tile_addr + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128) + ((unit ^ (uint32_t)(r & 7)) << 4);
                    if (!dirty_row) {                                   // no winner on this position: a zero operand row
                        if (lact) { sts128(addr, 0u, 0u, 0u, 0u); sts128(addr + BW_MAT, 0u, 0u, 0u, 0u); }
                    } else {
                        const uint32_t mb = DELTA ? 0xffu : (uint32_t)((m8 >> (8 * cur)) & 0xffull);   // delta: masks applied per entry
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float s0 = (mb >> (2 * q)) & 1u ? adj_scale : 0.f, s1 = (mb >> (2 * q + 1)) & 1u ? adj_scale : 0.f;
                            const float2 x = make_float2(acc[2 * q] * s0, acc[2 * q + 1] * s1);
                            const float2 h = make_float2(h_trunc(x.x), h_trunc(x.y));
                            const float2 l = sub2(x, h);
                            hi[q] = pack_h2(h.x, h.y);
                            lo[q] = pack_h2(l.x, l.y);
                        }
                        if (lact) {
                            sts128(addr, hi[0], hi[1], hi[2], hi[3]);
                            sts128(addr + BW_MAT, lo[0], lo[1], lo[2], lo[3]);
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                        dirty_row = false;
                    }
                    ++cur;
                    rend = lds_u16(rs + 2u * (uint32_t)min(p0 + r0 + cur + 1, P));
                };
                for (; e < eB; e += 4) {
                    float4 w[4][2];
                    float dj[4];
                    bool sd[4];
                    int jn[4];
#pragma unroll
                    for (int v = 0; v < 4; ++v) jn[v] = (e + v < eB) ? lds_u16(ls + 2u * (uint32_t)(e + v)) : 0;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        sd[v] = false;
                        if (DELTA) { sd[v] = (jn[v] >> 15) != 0; jn[v] &= 0x7FFF; }
                        w[v][0] = w[v][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (e + v < eB && lact) {
                            const float4* src = reinterpret_cast<const float4*>(wbase + (size_t)jn[v] * prm.kpad);
                            w[v][0] = __ldg(src);
                            w[v][1] = __ldg(src + 1);
                        }
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float d0 = lds_f32(dj_a + 4u * (uint32_t)jn[v]);
                        dj[v] = sd[v] ? -d0 : d0;
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        if (e + v < eB) {
                            while (e + v >= rend) store_row();
                            dirty_row = true;
                            const float dd = dj[v];
                            if (DELTA) {                               // relu mask of the entry's side, per channel
                                const uint32_t mb = (uint32_t)(((sd[v] ? m8x : m8) >> (8 * cur)) & 0xffull);
                                w[v][0].x = (mb & 1u) ? w[v][0].x : 0.f;   w[v][0].y = (mb & 2u) ? w[v][0].y : 0.f;
                                w[v][0].z = (mb & 4u) ? w[v][0].z : 0.f;   w[v][0].w = (mb & 8u) ? w[v][0].w : 0.f;
                                w[v][1].x = (mb & 16u) ? w[v][1].x : 0.f;  w[v][1].y = (mb & 32u) ? w[v][1].y : 0.f;
                                w[v][1].z = (mb & 64u) ? w[v][1].z : 0.f;  w[v][1].w = (mb & 128u) ? w[v][1].w : 0.f;
                            }
                            acc[0] = fmaf(dd, w[v][0].x, acc[0]); acc[1] = fmaf(dd, w[v][0].y, acc[1]);
                            acc[2] = fmaf(dd, w[v][0].z, acc[2]); acc[3] = fmaf(dd, w[v][0].w, acc[3]);
                            acc[4] = fmaf(dd, w[v][1].x, acc[4]); acc[5] = fmaf(dd, w[v][1].y, acc[5]);
                            acc[6] = fmaf(dd, w[v][1].z, acc[6]); acc[7] = fmaf(dd, w[v][1].w, acc[7]);
                        }
                    }
                }
                while (cur < 8) store_row();
                if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[tb]);
                if (PROF) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
            }
            if (tn == tpc - 1 && lane == 0) mbar_arrive(&recempty[rb]);   // past the chain's last tile: done with its record
            if (++tn == tpc) { tn = 0; ++ci; }
        }
        if (PROF && prm.prof && lane == 0 && (pw == 0 || pw == 15)) {
            long long* o = prm.prof + (size_t)blockIdx.x * 16 + (pw == 0 ? 8 : 12); o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; o[3] = pc[3];
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == BW_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// =====================================================================================================
// COMPACT delta backward.  The change of the adjoint rows between the current state and the proposal is non-zero on a few
// dozen positions only (cnn_winner_delta_kernel, compact record).  Instead of four mostly-zero 64-position tiles per
// (chain, net) the kernel builds ceil(npos / BD_NT) tiles - normally ONE - whose columns are the touched positions, and the
// epilogue turns the product Y[(t,a), c] into the SPARSE change of the per-net gradient
//     dGc[row, a] = sum over the columns c with pos[c] = row - t, t = 0..4, of Y[(t,a), c]      (ascending c: deterministic)
// for the output rows the record lists for the tile: the accumulator goes TMEM -> registers -> shared memory [c][(a,t)]
// (conflict-free, one store per column) and one thread per (row, a) gathers its <= 5 terms and writes the value straight
// to the scratch (coalesced); pas_reverse_accept_pos_kernel (or cnn_grad_combine_sparse_kernel) adds the three nets' lists to the
// row.  (The first version
// scatter-added column by column into a dense shared-memory row, one warp barrier per column, and flushed / re-zeroed 19 KB per
// chain and net: 6,100 of the ~9,500 cycles per tile, with the record buffers released only afterwards.)
// Roles and barriers as in cnn_backward_tc_kernel; differences: every role reads npos from the chain's record (the epilogue
// releases the record buffers too, it needs the lists), the global tile counter advances by the chain's own tile count,
// tiles have BD_NT = 48 columns, BD_NSET producer sets of BD_NW warps with BD_RPW columns each (3 x 6 x 8), 23 warps in all.
constexpr int BD_NREC_MAX = 6;          // record buffers: as many as fit in shared memory (BwdParams.nrec, >= 3)
constexpr int BD_MAT = BD_NT * KCH * 2; // one [48 x 64] fp16 operand matrix (6 KB)
constexpr int BD_SLOT = 2 * BD_MAT;     // hi + lo
constexpr int BD_NBUF = 3;             // operand tile buffers (3 x 48 KB).  4 buffers leave room for only 3 record buffers and
                                       // measured slower (7.1 vs 5.9 ms / 64k chains): the records are what lets the roles run ahead
constexpr int BD_TS = 100;              // floats per column of the transposed accumulator tile: index t * 20 + a
template <bool PROF>
__global__ void __launch_bounds__(BD_NTHREADS, 1) cnn_backward_delta_kernel(const __grid_constant__ BwdParams prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int C = prm.m.C, P = prm.m.P, J2 = 2 * C;
    const int nch = prm.nch;
    const int BD_NREC = prm.nrec;
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* sT = reinterpret_cast<float*>(ring + BD_NBUF * BW_MAXCH * BD_SLOT);    // [BD_NT][BD_TS] transposed accumulator tile
    float* sDj = sT + BD_NT * BD_TS;                                          // [J2] decoder weights
    uint16_t* sRec = reinterpret_cast<uint16_t*>((reinterpret_cast<uintptr_t>(sDj + J2) + 15) & ~(uintptr_t)15);   // [BD_NREC][rec]
    uint64_t* bars = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(sRec + BD_NREC * prm.rec) + 7) & ~(uintptr_t)7);
    uint64_t* full = bars;                      // [BD_NBUF] tile buffers: producers -> MMA (BD_NW warp arrivals: one producer set)
    uint64_t* empty = full + BD_NBUF;           // [BD_NBUF] MMA -> producers
    uint64_t* dfull = empty + BD_NBUF;          // [BW_NDBUF] MMA -> epilogue
    uint64_t* dempty = dfull + BW_NDBUF;        // [BW_NDBUF] epilogue -> MMA (4 warp arrivals)
    uint64_t* recfull = dempty + BW_NDBUF;      // [BD_NREC] record landed (bulk copy, tx bytes)
    uint64_t* recempty = recfull + BD_NREC_MAX; // [BD_NREC] producers and epilogue done with the record (18 + 4 warp arrivals)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(recempty + BD_NREC_MAX);

    const int k = blockIdx.x / prm.ctas_per_net;                       // this CTA's net
    const int within = blockIdx.x - k * prm.ctas_per_net;
    if (k >= prm.m.n_nets) return;
    const int b_lo = (int)((int64_t)prm.n * within / prm.ctas_per_net);
    const int b_hi = (int)((int64_t)prm.n * (within + 1) / prm.ctas_per_net);
    const int nchains = b_hi - b_lo;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const ppde_cnn_net_t net = prm.m.net[k];
    const uint32_t rec_bytes = (uint32_t)prm.rec * 2u;
    const uint32_t rec_a = smem_u32(sRec);

    if (threadIdx.x == 0) {
        for (int s = 0; s < BD_NBUF; ++s) { mbar_init(&full[s], BD_NW); mbar_init(&empty[s], 1); }
        for (int s = 0; s < BD_NREC; ++s) { mbar_init(&recfull[s], 1); mbar_init(&recempty[s], BD_NSET * BD_NW + NT_EPI / 32); }
        for (int d = 0; d < BW_NDBUF; ++d) { mbar_init(&dfull[d], 1); mbar_init(&dempty[d], NT_EPI / 32); }
        fence_barrier_init();
    }
    if (warp == BD_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int j = threadIdx.x; j < J2; j += BD_NTHREADS) sDj[j] = net.d[j] * net.adj_scale;   // (power of two: the operand rows come out scaled, exactly)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {   // A = W0^T (scaled, fp16 hi/lo) -> TMEM: lane m(a,t) = 32*(a/6) + 5*(a%6) + t
        const int grp = lane / 5, t = lane - 5 * grp, a = 6 * warp + grp;
        const bool rowok = lane < 30 && a < PPDE_Q;
        const int nrow = t * PPDE_Q + a;                                // W0r[c][t][a]
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int ks = 0; ks < prm.kpad / 16; ++ks) {
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int c0 = ks * 16 + 2 * q;
                const float w0 = (rowok && c0 < C) ? net.W0r[(size_t)c0 * 100 + nrow] * net.w0_scale : 0.f;
                const float w1 = (rowok && c0 + 1 < C) ? net.W0r[(size_t)(c0 + 1) * 100 + nrow] * net.w0_scale : 0.f;
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
        // ===== EPILOGUE: TMEM lane = (a, t) -> sT[c][a * 5 + t]; then one thread per (output row, a) gathers its <= 5 terms =====
        const int grp = lane / 5, d = lane - 5 * grp, a = 6 * warp + grp;
        const bool rowok = lane < 30 && a < PPDE_Q;
        const int tid = threadIdx.x;
        const float unscale = 1.f / (net.w0_scale * net.adj_scale);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16) + D_COL0;
        const uint32_t st_a = smem_u32(sT) + 4u * (uint32_t)(d * PPDE_Q + a);   // my element of column 0: sT[c][t * 20 + a]
        const uint32_t t_a = smem_u32(sT);
        long long pc[4] = {0, 0, 0, 0};
        long long tp = PROF ? clock64() : 0;
        int it = 0;
        int rb = -1;                                   // record buffer of the chain and the parity of its landing (BD_NREC is a
        uint32_t rph = 1u;                             // run-time value: `ci % BD_NREC` / `ci / BD_NREC` were real divisions, per chain and warp)
        for (int ci = 0; ci < nchains; ++ci) {
            if (++rb == BD_NREC) rb = 0;
            if (rb == 0) rph ^= 1u;
            mbar_wait(&recfull[rb], rph);
            const uint32_t rs = rec_a + (uint32_t)rb * rec_bytes;
            const int npos = lds_u16(rs);
            const int tiles = (npos + BD_NT - 1) / BD_NT;
            const uint32_t oo = rs + 2u * (uint32_t)(prm.rec - BD_MC * 32 - BD_TRAIL);   // trailer: ntile | nent | tstart[1 ..]
            const uint32_t pairs_a = rs + 2u * (uint32_t)(prm.rec - BD_MC * 32 - BD_TRAIL - 2 * (prm.vcap / PPDE_Q));   // fixed place
            float* vout = prm.Gc + ((size_t)k * prm.n + (b_lo + ci)) * prm.vcap;
            if (PROF) { const long long t1 = clock64(); pc[3] += t1 - tp; tp = t1; }
            for (int t = 0; t < tiles; ++t, ++it) {
                const int buf = it & (BW_NDBUF - 1);
                mbar_wait(&dfull[buf], (uint32_t)((it / BW_NDBUF) & 1));
                if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
                tc_fence_after();
                uint32_t y[BD_NT];
                tmem_ld32(lane_addr + buf * 64, y);
                tmem_ld16(lane_addr + buf * 64 + 32, y + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&dempty[buf]);
                if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                const int ncol = min(BD_NT, npos - t * BD_NT);
                named_bar(2, NT_EPI);                                      // the previous tile's gather has read sT
                if (rowok) {       // accumulator column j = operand row j holds column c = (j % BD_RPW) * 8 + j / BD_RPW (producers' permutation)
#pragma unroll
                    for (int j = 0; j < BD_NT; ++j) {
                        const int c = (j % BD_RPW) * BD_NW + j / BD_RPW;
                        if (c < ncol) sts_f32(st_a + (uint32_t)(c * BD_TS * 4), __uint_as_float(y[j]) * unscale);
                    }
                }
                named_bar(2, NT_EPI);
                // output rows of this tile: r in [tstart[t], tstart[t+1]); row i = orow[r] gets the columns c = cfirst[r] .. while
                // pos[c] <= i (at most 5, tap t = i - pos[c]); one thread per row: 5 x 16-byte loads per term, 80 bytes out
                const int r0 = t ? lds_u16(oo + 2u * (uint32_t)(1 + t)) : 0, r1 = lds_u16(oo + 2u * (uint32_t)(2 + t));   // tstart[t], tstart[t + 1]
                const uint32_t pa = rs + 2u * (uint32_t)(1 + t * BD_NT);                  // pos[] of the tile's columns
                for (int r = r0 + tid; r < r1; r += NT_EPI) {
                    const int i = lds_u16(pairs_a + 4u * (uint32_t)r);
                    const int cf = lds_u16(pairs_a + 4u * (uint32_t)r + 2u);
                    int pp[5];
#pragma unroll
                    for (int u = 0; u < 5; ++u) pp[u] = lds_u16(pa + 2u * (uint32_t)min(cf + u, ncol - 1));
                    float4 acc[5];
#pragma unroll
                    for (int q = 0; q < 5; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 5; ++u) {                                       // ascending columns: the order of the sums is fixed
                        const bool on = (cf + u < ncol) && (pp[u] <= i);                  // (pos >= i - 4 for every column >= cfirst)
                        const uint32_t src = t_a + 4u * (uint32_t)(min(cf + u, ncol - 1) * BD_TS + (on ? (i - pp[u]) : 0) * PPDE_Q);
#pragma unroll
                        for (int q = 0; q < 5; ++q) {
                            float4 v;
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src + 16u * q));
                            if (on) { acc[q].x += v.x; acc[q].y += v.y; acc[q].z += v.z; acc[q].w += v.w; }
                        }
                    }
                    float4* dst = reinterpret_cast<float4*>(vout + (size_t)r * PPDE_Q);
#pragma unroll
                    for (int q = 0; q < 5; ++q) __stcs(dst + q, acc[q]);
                }
                if (PROF) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&recempty[rb]);
        }
        if (PROF && prm.prof && threadIdx.x == 0) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; o[3] = pc[3]; }
    } else if (warp == BD_WARP_MMA) {
        // ===== MMA ISSUER: warp-uniform loop, one elected lane issues; also stages the records (bulk copies) =====
        const uint32_t idesc = make_idesc(128, BD_NT);
        const uint32_t ring_addr = smem_u32(ring);
        const int last_ksteps = (prm.kpad - (nch - 1) * KCH) / 16;
        const uint32_t a_lo_off = (uint32_t)(prm.kpad / 2);
        long long pc[3] = {0, 0, 0};
        long long tp = PROF ? clock64() : 0;
        // Records are staged up to BD_NREC chains ahead, OPPORTUNISTICALLY: buffer c % BD_NREC is free once the producers and
        // the epilogue have released chain c - BD_NREC; blocking on that here would serialise the MMAs of a chain behind the
        // epilogue of an earlier one, so the release is only polled (and waited for when the record is needed right now).
        int staged = 0, srb = 0;               // next chain to stage and its ring buffer
        uint32_t sph = 1u;                     // parity of the release that frees buffer srb for chain `staged` >= BD_NREC (toggles at
                                               // every wrap of srb: 0 for the first reuse)
        int nread = 0;                         // chains whose npos THIS warp has read: their buffers may be reused, not before
                                               // (a chain without tiles is released by the other roles without waiting for the MMAs)
        auto stage_records = [&](int need) {   // stage what is free; block until chains < need are staged
            while (staged < nchains) {
                const int c = staged;
                if (c >= BD_NREC) {
                    if (c - BD_NREC >= nread) break;
                    if (c < need) mbar_wait(&recempty[srb], sph);
                    else if (!mbar_test(&recempty[srb], sph)) break;
                }
                if (elect_one()) {
                    mbar_expect_tx(&recfull[srb], rec_bytes);
                    bulk_g2s(sRec + (size_t)srb * prm.rec, prm.wl + ((size_t)(b_lo + c) * prm.m.n_nets + k) * prm.rec, rec_bytes,
                             &recfull[srb]);
                }
                __syncwarp();
                ++staged;
                if (++srb == BD_NREC) { srb = 0; sph ^= 1u; }
            }
        };
        int it = 0;
        int rb = -1;
        uint32_t rph = 1u;
        for (int ci = 0; ci < nchains; ++ci) {
            stage_records(ci + 1);
            if (++rb == BD_NREC) rb = 0;
            if (rb == 0) rph ^= 1u;
            mbar_wait(&recfull[rb], rph);
            const int npos = lds_u16(rec_a + (uint32_t)rb * rec_bytes);
            nread = ci + 1;
            const int tiles = (npos + BD_NT - 1) / BD_NT;
            for (int t = 0; t < tiles; ++t, ++it) {
                const int tb = it % BD_NBUF;
                const int buf = it & (BW_NDBUF - 1);
                if (it >= BW_NDBUF) mbar_wait(&dempty[buf], (uint32_t)(((it / BW_NDBUF) + 1) & 1));
                if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
                // keep the records flowing while the tile is being built: bounded SUSPENDING waits, not a spin (this is the
                // top-priority warp of its scheduler: spinning here starves four producer warps, measured as a 10x slower kernel)
                while (!mbar_try_wait_for(&full[tb], (uint32_t)((it / BD_NBUF) & 1), 2000u)) stage_records(0);
                if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + D_COL0 + buf * 64;
                if (elect_one()) {
                    for (int kc = 0; kc < nch; ++kc) {
                        const int slot = tb * BW_MAXCH + kc;
                        const uint64_t dhi = make_b_desc(ring_addr + slot * BD_SLOT);
                        const uint64_t dlo = make_b_desc(ring_addr + slot * BD_SLOT + BD_MAT);
                        const int ksteps = (kc == nch - 1) ? last_ksteps : KCH / 16;
                        const uint32_t a_hi0 = tmem_base + kc * (KCH / 2);
#pragma unroll
                        for (int ks = 0; ks < KCH / 16; ++ks) {
                            if (ks < ksteps) {
                                const uint32_t a_hi = a_hi0 + ks * 8;
                                mma_ts(d_tmem, a_hi, dhi + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
                                mma_ts(d_tmem, a_hi, dlo + (uint64_t)(ks * 2), idesc, 1u);
                                mma_ts(d_tmem, a_hi + a_lo_off, dhi + (uint64_t)(ks * 2), idesc, 1u);
                            }
                        }
                    }
                    tc_commit(&empty[tb]);
                    tc_commit(&dfull[buf]);
                }
                __syncwarp();
                if (PROF) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
                stage_records(0);
            }
        }
        if (PROF && prm.prof && lane == 0) { long long* o = prm.prof + (size_t)blockIdx.x * 16; o[4] = pc[0]; o[5] = pc[1]; o[6] = pc[2]; o[7] = it; }
    } else {
        // ===== PRODUCERS: BD_NSET sets of BD_NW warps take the tiles in turn (set = it % BD_NSET).  Warp w of a set owns the columns
        // c = w, w + BD_NW, w + 2 BD_NW, .. of the tile (STRIDED: the touched positions come in runs - the 5 conv rows of a mutated
        // residue carry every winner sitting on them, both sides, ~4 entries per column, while a moved winner's column carries one -
        // and with contiguous ownership one warp got a whole run, ~2.5x the mean, and the tile waited for it).  The operand row of
        // column c is r = (c % BD_NW) * BD_RPW + c / BD_NW, so a warp's rows are contiguous; the epilogue undoes the permutation when
        // it writes sT.  The record lists the entries in exactly this order (cnn_delta_record_kernel): a warp reads the two offsets of
        // its (tile, warp) group and walks the list front to back, four W1 rows in flight across column boundaries. =====
        const int pw = warp - 4;
        const int pset = pw / BD_NW, w8 = pw - pset * BD_NW;      // set, warp of the set
        const bool lact = 8 * lane < prm.kpad;
        const float* wbase = net.W1p + 8 * lane;
        const uint32_t ring_lane = smem_u32(ring) + (uint32_t)((lane >> 3) * BD_SLOT);
        const uint32_t unit = (uint32_t)(lane & 7);
        const uint32_t dj_a = smem_u32(sDj);
        const int NB = prm.NB;
        long long pc[4] = {0, 0, 0, 0};
        long long tp = PROF ? clock64() : 0;
        int it = 0;
        int rbp = (BD_NREC - 1) | 16;                  // record buffer (low 4 bits) and landing parity (bit 4) of the chain, one register
        for (int ci = 0; ci < nchains; ++ci) {
            rbp = ((rbp & 15) + 1 == BD_NREC) ? ((rbp & 16) ^ 16) : rbp + 1;
            const int rb = rbp & 15;
            const uint32_t rph = (uint32_t)(rbp >> 4);
            const int bb = b_lo + ci;
            if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
            mbar_wait(&recfull[rb], rph);
            if (PROF) { const long long t1 = clock64(); pc[3] += t1 - tp; tp = t1; }
            const uint32_t rs = rec_a + (uint32_t)rb * rec_bytes;
            const int npos = lds_u16(rs);
            const int tiles = (npos + BD_NT - 1) / BD_NT;
            const uint32_t ps = rs + 2u;                                   // pos[c]
            const uint32_t ws = rs + 2u * (uint32_t)(1 + npos);            // woff[8 t + w]: my entries of tile t are list[woff .. woff')
            const uint32_t ls = ws + 2u * (uint32_t)(BD_NW * tiles + 1);   // list, in the producers' work order (cnn_delta_record_kernel)
            for (int t = 0; t < tiles; ++t, ++it) {
                if (it % BD_NSET != pset) continue;
                const int tb = it % BD_NBUF;
                const uint32_t tparity = (uint32_t)(((it / BD_NBUF) + 1) & 1);
                const int ncol = min(BD_NT, npos - t * BD_NT);
                const int cbase = t * BD_NT + w8;                          // my columns: cbase + 8 i, i < BD_RPW, while < t * BD_NT + ncol
                const int f_lo = lds_u16(ws + 2u * (uint32_t)(BD_NW * t + w8));
                const int ntot = lds_u16(ws + 2u * (uint32_t)(BD_NW * t + w8 + 1)) - f_lo;
                // relu-mask bytes of my columns, both sides (byte i of the 64-bit words = column slot i)
                unsigned long long m8 = 0ull, m8x = 0ull;
#pragma unroll
                for (int i = 0; i < BD_RPW; ++i) {
                    const int c = cbase + BD_NW * i;
                    if (w8 + BD_NW * i < ncol && lact) {
                        if (c < BD_MC) {                    // from the record (shared memory)
                            const uint32_t ma = rs + rec_bytes - (uint32_t)(BD_MC * 64) + (uint32_t)(c * 64) + (uint32_t)lane;
                            uint32_t by, bx;
                            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(by) : "r"(ma));
                            asm volatile("ld.shared.u8 %0, [%1+32];" : "=r"(bx) : "r"(ma));
                            m8 |= (unsigned long long)by << (8 * i);
                            m8x |= (unsigned long long)bx << (8 * i);
                        } else {                            // beyond the record's capacity: block table -> mask row in global memory
                            const int mry = __ldg(prm.mask_rows + bb), mrx = __ldg(prm.mask_rows_x + bb);
                            const int p = lds_u16(ps + 2u * (uint32_t)c);
                            int ry = mry, rx = mrx;
                            if (prm.btab) { ry = __ldg(prm.btab + (size_t)mry * NB + (p >> PB_SHIFT)); rx = __ldg(prm.btab + (size_t)mrx * NB + (p >> PB_SHIFT)); }
                            m8 |= (unsigned long long)__ldg(prm.r1mask + (((size_t)ry * prm.m.n_nets + k) * P + p) * 32 + lane) << (8 * i);
                            m8x |= (unsigned long long)__ldg(prm.r1mask + (((size_t)rx * prm.m.n_nets + k) * P + p) * 32 + lane) << (8 * i);
                        }
                    }
                }
                const uint32_t tile_addr = ring_lane + (uint32_t)(tb * BW_MAXCH * BD_SLOT);
                int cur = 0;                                // column slot being accumulated (rows before it are stored)
                bool have_buf = false, dirty_row = false;
                float acc[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                auto store_row = [&]() {   // operand row of slot `cur` <- scale * acc (masks were applied per entry), fp16 hi + lo
                    if (!have_buf) {
                        if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                        mbar_wait(&empty[tb], tparity);
                        have_buf = true;
                        if (PROF) { const long long t1 = clock64(); pc[0] += t1 - tp; tp = t1; }
                    }
                    const int r = w8 * BD_RPW + cur;
                    const uint32_t addr = tile_addr + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128) + ((unit ^ (uint32_t)(r & 7)) << 4);
                    if (!dirty_row) {
                        if (lact) { sts128(addr, 0u, 0u, 0u, 0u); sts128(addr + BD_MAT, 0u, 0u, 0u, 0u); }
                    } else {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 x = make_float2(acc[2 * q], acc[2 * q + 1]);
                            const float2 h = make_float2(h_trunc(x.x), h_trunc(x.y));
                            const float2 l = sub2(x, h);
                            hi[q] = pack_h2(h.x, h.y);
                            lo[q] = pack_h2(l.x, l.y);
                        }
                        if (lact) {
                            sts128(addr, hi[0], hi[1], hi[2], hi[3]);
                            sts128(addr + BD_MAT, lo[0], lo[1], lo[2], lo[3]);
                        }
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                        dirty_row = false;
                    }
                    ++cur;
                };
                // my entries, front to back: entry = channel | slot << 12 | side << 15, four W1 rows in flight across column boundaries
                const uint32_t fl = ls + 2u * (uint32_t)f_lo;
                for (int e = 0; e < ntot; e += 4) {
                    float4 w[4][2];
                    float dj[4];
                    uint32_t fe[4];                         // the entries stay packed: slot / side / channel are decoded where they are used
#pragma unroll
                    for (int v = 0; v < 4; ++v) fe[v] = (e + v < ntot) ? (uint32_t)lds_u16(fl + 2u * (uint32_t)(e + v)) : 0u;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        w[v][0] = w[v][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (e + v < ntot && lact) {
                            const float4* src = reinterpret_cast<const float4*>(wbase + (size_t)(fe[v] & 0x1FFu) * prm.kpad);
                            w[v][0] = __ldg(src);
                            w[v][1] = __ldg(src + 1);
                        }
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float d0 = lds_f32(dj_a + 4u * (fe[v] & 0x1FFu));
                        dj[v] = (fe[v] >> 15) ? -d0 : d0;
                    }
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        if (e + v < ntot) {
                            const int slv = (int)((fe[v] >> 12) & 7u);
                            while (cur < slv) store_row();
                            dirty_row = true;
                            const float dd = dj[v];
                            const uint32_t mb = (uint32_t)((((fe[v] >> 15) ? m8x : m8) >> (8 * cur)) & 0xffull);   // relu mask of the entry's side
                            // (a masked-off channel contributes exactly 0: skipping the FMA leaves acc unchanged, same value)
                            if (mb & 1u) acc[0] = fmaf(dd, w[v][0].x, acc[0]);
                            if (mb & 2u) acc[1] = fmaf(dd, w[v][0].y, acc[1]);
                            if (mb & 4u) acc[2] = fmaf(dd, w[v][0].z, acc[2]);
                            if (mb & 8u) acc[3] = fmaf(dd, w[v][0].w, acc[3]);
                            if (mb & 16u) acc[4] = fmaf(dd, w[v][1].x, acc[4]);
                            if (mb & 32u) acc[5] = fmaf(dd, w[v][1].y, acc[5]);
                            if (mb & 64u) acc[6] = fmaf(dd, w[v][1].z, acc[6]);
                            if (mb & 128u) acc[7] = fmaf(dd, w[v][1].w, acc[7]);
                        }
                    }
                }
                while (cur < BD_RPW) store_row();
                if (PROF) { const long long t1 = clock64(); pc[1] += t1 - tp; tp = t1; }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[tb]);
                if (PROF) { const long long t1 = clock64(); pc[2] += t1 - tp; tp = t1; }
            }
            if (lane == 0) mbar_arrive(&recempty[rb]);                     // done with this chain's record
        }
        if (PROF && prm.prof && lane == 0 && (pw == 0 || pw == 15)) {
            long long* o = prm.prof + (size_t)blockIdx.x * 16 + (pw == 0 ? 8 : 12); o[0] = pc[0]; o[1] = pc[1]; o[2] = pc[2]; o[3] = pc[3];
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == BD_WARP_MMA) {
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

extern "C" int ppde_cnn_forward_tc(const ppde_cnn_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n,
                                   unsigned long long* mkey, uint8_t* r1mask, const ppde_tune_t* tune, void* stream) {
    if (n <= 0) return 0;
    if (!m || !aa || !mkey || aa_stride < m->L) return (int)cudaErrorInvalidValue;
    if (m->C > 256 || m->P < 1) return (int)cudaErrorInvalidValue;       // A must fit 256 TMEM columns
    const int forward_variant = (tune && tune->forward_ctas == 1) ? 1 : 2;
    long long* const forward_prof = tune ? tune->prof : nullptr;
    tc::Params prm;
    prm.m = *m;
    prm.aa = aa;
    prm.aa_stride = aa_stride;
    prm.n = n;
    prm.mkey = mkey;
    prm.r1mask = r1mask;
    prm.prof = nullptr;
    prm.kpad = (m->C + 15) / 16 * 16;
    prm.nch = (prm.kpad + tc::KCH - 1) / tc::KCH;
    const int sms = sm_count();
    const int MT = (2 * m->C + 127) / 128;
    cudaStream_t st = (cudaStream_t)stream;
    if (forward_variant == 2) {
        prm.n_tile = tc::NT2;
        prm.tiles_per_chain = (m->P + tc::NT2 - 1) / tc::NT2;
        prm.MT = (MT + 1) / 2;                                             // channel-tile pairs
        const int combos = m->n_nets * prm.MT;
        prm.ctas_per_combo = (sms / 2) / combos;                           // cluster pairs per combo
        if (prm.ctas_per_combo < 1) prm.ctas_per_combo = 1;
        if (prm.ctas_per_combo > n) prm.ctas_per_combo = n;
        const size_t smem = (size_t)tc::NSLOT2 * tc::SLOT2_BYTES + (size_t)100 * prm.nch * tc::KCH * sizeof(float) +
                            32 * sizeof(uint64_t) + 1024;
        void (*kern)(tc::Params) = nullptr;
        switch (prm.nch) {
            case 1: kern = tc::cnn_forward_tc2_kernel<1, false>; break;
            case 2: kern = tc::cnn_forward_tc2_kernel<2, false>; break;
            case 3: kern = tc::cnn_forward_tc2_kernel<3, false>; break;
            default: kern = tc::cnn_forward_tc2_kernel<4, false>; break;
        }
        prm.prof = forward_prof;
        if (forward_prof && prm.nch == 4) kern = tc::cnn_forward_tc2_kernel<4, true>;   // role-level cycle counters (tools/prof_fwd.py)
        static SmemCache configured2[5];
        const int cfg = (forward_prof && prm.nch == 4) ? 0 : prm.nch;
        if (cudaError_t e = ensure_dynamic_smem(kern, smem, configured2[cfg])) return (int)e;
        kern<<<2 * combos * prm.ctas_per_combo, tc::NTHREADS, smem, st>>>(prm);
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
    static SmemCache configured;
    if (cudaError_t e = ensure_dynamic_smem(tc::cnn_forward_tc_kernel, smem, configured)) return (int)e;
    tc::cnn_forward_tc_kernel<<<combos * prm.ctas_per_combo, tc::NTHREADS, smem, st>>>(prm);
    return launch_done();
}

extern "C" int ppde_cnn_dirty(const ppde_cnn_t* m, const uint8_t* aa_x, const uint8_t* aa_y, int32_t aa_stride, int32_t n,
                              uint32_t* dmask, void* stream) {
    if (n <= 0) return 0;
    if (!aa_x || !aa_y || !dmask) return (int)cudaErrorInvalidValue;
    tc::cnn_dirty_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(n, m->L, m->P, aa_stride, aa_x, aa_y, dmask);
    return launch_done();
}

extern "C" int64_t ppde_cnn_forward_inc_ws_bytes(int32_t n) { return ((int64_t)n + 256 + (int64_t)n * tc::PB_MAXNB) * 4; }
extern "C" int32_t ppde_cnn_block_positions(void) { return tc::PB; }

extern "C" int ppde_cnn_forward_inc(const ppde_cnn_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n,
                                    unsigned long long* mkey, uint8_t* r1mask, const uint32_t* dmask,
                                    unsigned long long* bkey, int32_t* btab, const int32_t* rows_x, const int32_t* rows_y,
                                    int32_t row_base_y, unsigned long long* mkey_pool, void* ws, const ppde_tune_t* tune,
                                    void* stream) {
    if (n <= 0) return 0;
    if (!m || !aa || aa_stride < m->L) return (int)cudaErrorInvalidValue;
    const int NB = (m->P + tc::PB - 1) / tc::PB;
    if (m->C > 256 || m->P < 1 || NB > tc::PB_MAXNB || !bkey || !btab || !mkey || !ws || (dmask && !rows_x)) return (int)cudaErrorInvalidValue;
    const int inc_parts = (tune && (tune->parts & 7)) ? (tune->parts & 7) : 7;
    tc::IncParams prm;
    prm.m = *m; prm.aa = aa; prm.aa_stride = aa_stride; prm.n = n; prm.mkey = mkey; prm.r1mask = r1mask; prm.dmask = dmask;
    prm.bkey = bkey; prm.btab = btab; prm.mkey_pool = mkey_pool;
    prm.rows_x = dmask ? rows_x : nullptr;   // full evaluation: nothing is read from a current row
    prm.rows_y = rows_y; prm.row_base_y = row_base_y; prm.NB = NB;
    prm.kpad = (m->C + 15) / 16 * 16;
    prm.dbg = tune ? tune->dbg : 0;
    prm.prof = tune ? tune->prof : nullptr;
    const int nch = (prm.kpad + tc::KCH - 1) / tc::KCH;
    const int sms = sm_count();
    const int MT = (2 * m->C + 127) / 128;
    prm.MT = (MT + 1) / 2;
    const int combos = m->n_nets * prm.MT;
    prm.ctas_per_combo = (sms / 2) / combos;
    if (prm.ctas_per_combo < 1) prm.ctas_per_combo = 1;
    if (prm.ctas_per_combo > n) prm.ctas_per_combo = n;
    if (prm.ctas_per_combo > 256) prm.ctas_per_combo = 256;
    int32_t* boff = reinterpret_cast<int32_t*>(ws);
    int32_t* gtot = boff + n;
    uint32_t* blist = reinterpret_cast<uint32_t*>(gtot + 256);
    prm.boff = boff; prm.gtot = gtot; prm.blist = blist;
    if (inc_parts & 1) {
        tc::cnn_inc_scan_kernel<<<prm.ctas_per_combo, 1024, 0, (cudaStream_t)stream>>>(n, prm.ctas_per_combo, NB, dmask, boff, gtot, blist);
        int r0 = launch_done();
        if (r0) return r0;
    }
    const size_t smem = (size_t)tc::NSLOT2 * tc::SLOT2_BYTES + (size_t)100 * nch * tc::KCH * sizeof(float) +
                        32 * sizeof(uint64_t) + 1024;
    void (*kern)(tc::IncParams) = nullptr;
    switch (nch) {
        case 1: kern = tc::cnn_forward_inc_kernel<1>; break;
        case 2: kern = tc::cnn_forward_inc_kernel<2>; break;
        case 3: kern = tc::cnn_forward_inc_kernel<3>; break;
        default: kern = tc::cnn_forward_inc_kernel<4>; break;
    }
    static SmemCache configured[5];
    if (cudaError_t e = ensure_dynamic_smem(kern, smem, configured[nch])) return (int)e;
    if (inc_parts & 2) {
        kern<<<2 * combos * prm.ctas_per_combo, tc::INC_NTHREADS, smem, (cudaStream_t)stream>>>(prm);
        int r1 = launch_done();
        if (r1) return r1;
    }
    if (inc_parts & 4) {
        tc::cnn_inc_merge_kernel<<<n, tc::MERGE_NT, 0, (cudaStream_t)stream>>>(prm);
        return launch_done();
    }
    return 0;
}

extern "C" int ppde_cnn_backward_tc_rows(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa, int32_t aa_stride,
                                         int32_t n, const unsigned long long* mkey, float lamda,
                                         const float* Gp, int64_t Gp_stride, const int32_t* gp_rows,
                                         float* G, int64_t G_stride, const int32_t* g_rows, const uint8_t* r1mask,
                                         const int32_t* mask_rows, int32_t mask_row_base, const int32_t* btab, float* scratch,
                                         const ppde_tune_t* tune, void* stream);
extern "C" int ppde_cnn_backward_tc(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa, int32_t aa_stride,
                                    int32_t n, const unsigned long long* mkey, float lamda,
                                    const float* Gp, int64_t Gp_stride, const int32_t* gp_rows,
                                    float* G, int64_t G_stride, const int32_t* g_rows, const uint8_t* r1mask,
                                    float* scratch, const ppde_tune_t* tune, void* stream) {
    return ppde_cnn_backward_tc_rows(m, pm, aa, aa_stride, n, mkey, lamda, Gp, Gp_stride, gp_rows, G, G_stride, g_rows, r1mask,
                                     nullptr, 0, nullptr, scratch, tune, stream);
}
struct BwdDelta {                 // delta backward: gradient of the proposal = gradient of the current state + change
    const uint8_t* aa_x;          // current states [n, aa_stride]
    const unsigned long long* mkey_pool;   // [rows, n_nets, 2C] winners of every pool row
    const int32_t* rows_x;        // [n] pool row of the current state (G, Gp, r1mask, mkey_pool)
    const int32_t* rows_y;        // [n] pool row of the proposal
};

static int backward_launch(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa, int32_t aa_stride,
                           int32_t n, const unsigned long long* mkey, float lamda,
                           const float* Gp, int64_t Gp_stride, const int32_t* gp_rows,
                           float* G, int64_t G_stride, const int32_t* g_rows, const uint8_t* r1mask,
                           const int32_t* mask_rows, int32_t mask_row_base, const int32_t* btab, float* scratch,
                           const ppde_tune_t* tune, void* stream, const BwdDelta* dl) {
    if (n <= 0) return 0;
    if (!m || !pm || !aa || !mkey || !G || aa_stride < m->L || (G_stride & 3) || (Gp && (Gp_stride & 3))) return (int)cudaErrorInvalidValue;
    if (m->C > 256 || m->P < 1 || !scratch || !r1mask) return (int)cudaErrorInvalidValue;
    const int bwd_parts = (tune && (tune->parts & 7)) ? (tune->parts & 7) : 7;
    tc::BwdParams prm;
    prm.m = *m; prm.pm = *pm; prm.aa = aa; prm.aa_stride = aa_stride; prm.n = n; prm.mkey = mkey;
    prm.Gc = scratch;
    prm.r1mask = r1mask;
    prm.mask_rows = dl ? dl->rows_y : mask_rows;
    prm.mask_row_base = mask_row_base;
    prm.mask_rows_x = dl ? dl->rows_x : nullptr;
    prm.btab = btab; prm.NB = (m->P + tc::PB - 1) / tc::PB;
    prm.dbg = tune ? tune->dbg : 0;
    prm.tiles_per_chain = (m->P + tc::BW_NT - 1) / tc::BW_NT;
    prm.kpad = (m->C + 15) / 16 * 16;
    prm.nch = (prm.kpad + tc::KCH - 1) / tc::KCH;
    const int sms = sm_count();
    prm.ctas_per_net = sms / m->n_nets;
    if (prm.ctas_per_net < 1) prm.ctas_per_net = 1;
    if (prm.ctas_per_net > n) prm.ctas_per_net = n;
    const int C = m->C, P = m->P, L = m->L, J2 = 2 * C;
    // delta mode, compact records (default; tune->delta_layout = 1 keeps one column per position): npos | pos | start | list
    const bool compact = dl && !(tune && tune->delta_layout == 1);
    // compact record: npos | pos[P] | woff[8 TMAX + 1] | list[2 J2] | ... | (orow, cfirst)[RMAX] | trailer[BD_TRAIL] | masks[BD_MC][2][32 B]  (uint16)
    const int tmax = (P + tc::BD_NT - 1) / tc::BD_NT, rmax = L + 4 * tmax;
    const int rec = compact ? (((1 + P + 8 * tmax + 1 + 2 * J2 + 2 * rmax + 7) & ~7) + tc::BD_TRAIL + tc::BD_MC * 32) : (((P + 1) + 2 * J2 + 7) & ~7);
    const int vcap = compact ? rmax * PPDE_Q : L * PPDE_Q;                       // floats of scratch per (net, chain)
    prm.vcap = vcap;
    const size_t smem_fixed = compact
        ? 1024 + (size_t)tc::BD_NBUF * tc::BW_MAXCH * tc::BD_SLOT + ((size_t)tc::BD_NT * tc::BD_TS + J2) * sizeof(float) + 16 + 8 + 32 * sizeof(uint64_t)
        : 1024 + (size_t)tc::BW_NBUF * tc::BW_MAXCH * tc::BW_SLOT + ((size_t)L * PPDE_Q + 4 + J2) * sizeof(float) + 16 + 8 + 32 * sizeof(uint64_t);
    int nrec = 2;
    if (compact) {                                   // as many record buffers as fit under the 227 KB limit (3 .. BD_NREC_MAX)
        nrec = (int)((232448 - smem_fixed) / ((size_t)rec * sizeof(uint16_t)));
        if (nrec > tc::BD_NREC_MAX) nrec = tc::BD_NREC_MAX;
        if (nrec < 3) return (int)cudaErrorInvalidValue;
    }
    prm.nrec = nrec;
    const size_t smem = smem_fixed + (size_t)nrec * rec * sizeof(uint16_t);
    long long* const backward_prof = tune ? tune->prof : nullptr;
    const bool prof = backward_prof != nullptr;
    void (*bkern)(tc::BwdParams) =
        compact ? (prof ? tc::cnn_backward_delta_kernel<true> : tc::cnn_backward_delta_kernel<false>)
                : (dl ? (prof ? tc::cnn_backward_tc_kernel<true, true> : tc::cnn_backward_tc_kernel<false, true>)
                      : (prof ? tc::cnn_backward_tc_kernel<true, false> : tc::cnn_backward_tc_kernel<false, false>));
    prm.prof = backward_prof;
    static SmemCache configured[6];
    const int cfg = (compact ? 4 : (dl ? 2 : 0)) + (prof ? 1 : 0);
    if (cudaError_t e = ensure_dynamic_smem(bkern, smem, configured[cfg])) return (int)e;
    cudaStream_t st = (cudaStream_t)stream;
    // winner records live behind the per-net gradient scratch: [n_nets*n*vcap floats][n*n_nets*rec uint16]
    uint16_t* wl = reinterpret_cast<uint16_t*>(scratch + (size_t)m->n_nets * n * vcap);
    prm.wl = wl;
    prm.rec = rec;
    if (bwd_parts & 1) {
        if (compact)
            tc::cnn_delta_record_kernel<<<n * m->n_nets, 128, (4 * J2 + P) * sizeof(int), st>>>(
                *m, m->n_nets, C, P, L, aa_stride, dl->aa_x, aa, mkey, dl->mkey_pool, dl->rows_x, wl, rec, r1mask, btab, prm.NB, dl->rows_y);
        else if (dl)
            tc::cnn_winner_delta_kernel<<<n * m->n_nets, 128, ((P + 1) + 2 * P + 4 * J2) * sizeof(int), st>>>(
                *m, m->n_nets, C, P, L, aa_stride, dl->aa_x, aa, mkey, dl->mkey_pool, dl->rows_x, wl, rec, 0,
                nullptr, btab, prm.NB, dl->rows_y);
        else
            tc::cnn_winner_sort_kernel<<<n * m->n_nets, 128, ((P + 1) + P + 2 * J2) * sizeof(int), st>>>(m->n_nets, C, P, mkey, wl, rec);
        int r0 = launch_done();
        if (r0) return r0;
    }
    if (bwd_parts & 2) {
        bkern<<<m->n_nets * prm.ctas_per_net, compact ? tc::BD_NTHREADS : tc::BW_NTHREADS, smem, st>>>(prm);
        int r = launch_done();
        if (r) return r;
    }
    if (bwd_parts & 4) {
        if (compact) {
            const size_t csm = (size_t)L * PPDE_Q * sizeof(float);
            static SmemCache cconf;
            if (cudaError_t e = ensure_dynamic_smem(tc::cnn_grad_combine_sparse_kernel, csm, cconf)) return (int)e;
            tc::cnn_grad_combine_sparse_kernel<<<n, 256, csm, st>>>(n, L, m->n_nets, lamda / (float)m->n_nets, *pm, scratch, vcap, wl, rec,
                                                                    Gp, Gp_stride, G, G_stride, dl->rows_x, dl->rows_y);
        } else if (dl)
            tc::cnn_grad_combine_delta_kernel<<<n, 256, 0, st>>>(n, L * PPDE_Q, m->n_nets, lamda / (float)m->n_nets, *pm, scratch,
                                                                 Gp, Gp_stride, G, G_stride, dl->rows_x, dl->rows_y);
        else
            tc::cnn_grad_combine_kernel<<<n, 256, 0, st>>>(n, L * PPDE_Q, m->n_nets, lamda / (float)m->n_nets, *pm, scratch,
                                                           Gp, Gp_stride, gp_rows, G, G_stride, g_rows);
        return launch_done();
    }
    return 0;
}

extern "C" int ppde_cnn_backward_delta_layout(const ppde_cnn_t* m, int32_t n, int32_t* vcap, int32_t* rec, int64_t* wl_offset) {
    if (!m || n <= 0 || !vcap || !rec || !wl_offset) return (int)cudaErrorInvalidValue;
    const int P = m->P, L = m->L, J2 = 2 * m->C;                 // same formulas as backward_launch (compact records)
    const int tmax = (P + tc::BD_NT - 1) / tc::BD_NT, rmax = L + 4 * tmax;
    *rec = ((1 + P + 8 * tmax + 1 + 2 * J2 + 2 * rmax + 7) & ~7) + tc::BD_TRAIL + tc::BD_MC * 32;
    *vcap = rmax * PPDE_Q;
    *wl_offset = (int64_t)m->n_nets * n * (*vcap);
    return 0;
}

// floats of `scratch` the tensor-core backward entry points need for n chains (the largest of the three layouts)
extern "C" int64_t ppde_cnn_backward_scratch_floats(const ppde_cnn_t* m, int32_t n) {
    if (!m || n <= 0) return 0;
    const int64_t P = m->P, L = m->L, J2 = 2 * (int64_t)m->C;
    const int64_t tmax = (P + tc::BD_NT - 1) / tc::BD_NT, rmax = L + 4 * tmax;
    const int64_t rec_c = ((1 + P + 8 * tmax + 1 + 2 * J2 + 2 * rmax + 7) & ~(int64_t)7) + tc::BD_TRAIL + tc::BD_MC * 32, rec_d = ((P + 1) + 2 * J2 + 7) & ~(int64_t)7;
    const int64_t a = (int64_t)m->n_nets * n * rmax * PPDE_Q + ((int64_t)n * m->n_nets * rec_c + 1) / 2;     // compact delta
    const int64_t b = (int64_t)m->n_nets * n * L * PPDE_Q + ((int64_t)n * m->n_nets * rec_d + 1) / 2;        // exact / per-position delta
    return (a > b ? a : b) + 16;
}

extern "C" int ppde_cnn_backward_tc_rows(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa, int32_t aa_stride,
                                         int32_t n, const unsigned long long* mkey, float lamda,
                                         const float* Gp, int64_t Gp_stride, const int32_t* gp_rows,
                                         float* G, int64_t G_stride, const int32_t* g_rows, const uint8_t* r1mask,
                                         const int32_t* mask_rows, int32_t mask_row_base, const int32_t* btab, float* scratch,
                                         const ppde_tune_t* tune, void* stream) {
    return backward_launch(m, pm, aa, aa_stride, n, mkey, lamda, Gp, Gp_stride, gp_rows, G, G_stride, g_rows, r1mask, mask_rows,
                           mask_row_base, btab, scratch, tune, stream, nullptr);
}

extern "C" int ppde_cnn_backward_delta(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa_x, const uint8_t* aa_y,
                                       int32_t aa_stride, int32_t n, const unsigned long long* mkey_y,
                                       const unsigned long long* mkey_pool, float lamda, const float* Gp, int64_t Gp_stride,
                                       float* G, int64_t G_stride, const int32_t* rows_x, const int32_t* rows_y,
                                       const uint8_t* r1mask, const int32_t* btab, float* scratch, const ppde_tune_t* tune,
                                       void* stream) {
    if (!aa_x || !mkey_pool || !rows_x || !rows_y || !G) return (int)cudaErrorInvalidValue;
    BwdDelta dl{aa_x, mkey_pool, rows_x, rows_y};
    return backward_launch(m, pm, aa_y, aa_stride, n, mkey_y, lamda, Gp, Gp_stride, rows_y, G, G_stride, rows_y, r1mask, rows_y, 0,
                           btab, scratch, tune, stream, &dl);
}
