// Dense Potts field on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a only.
//
//   Gp[b, (j,l)] = h[(j,l)] + sum_{(i,k)} X[b,(i,k)] * Jsym[(i,k),(j,l)]        X = one-hot of the window residues
//
// i.e. the reference's two einsums + autograd (PottsModel.hamiltonian, ppde/nets.py:282-290; ppde/energy.py:106-108)
// as ONE GEMM  [n x D] x [D x D]  per full re-evaluation, used instead of the row-gather
// kernel (potts.cu) when many chains are evaluated from scratch (t = 0, periodic refresh, the L-scaling sweep).
//
// GEMM view per CTA tile:   D[col, chain] = sum_k Jsym[col, k] * X[chain, k]       (Jsym is symmetric)
//   M = 128 field columns, N = 256 chains, K = D = 20 Lp in chunks of 64.
//   * A = Jsym tile, fp16 hi (truncated: exact) + lo (rounded residual) of Jsym * jscale (power of two), taken from a
//     PRE-TILED image in global memory (ppde_potts_dense_pack): every (column tile, K chunk) is one contiguous 32 KB
//     block already in the K-major SWIZZLE_128B shared-memory layout, so a stage is filled by ONE 1-D bulk copy (TMA
//     engine) with mbarrier transaction-byte completion.  X is exact in fp16, so hi + lo gives the field to ~2^-22
//     relative per term with fp32 accumulation in tensor memory (2 MMA passes per K step).
//   * B = one-hot tile [256 chains x 64], generated on the fly by 8 producer warps from the residue bytes (zero fill
//     + <= 4 ones per row), K-major SW128.
//   * D (128 x 256 fp32) is double-buffered in TMEM (512 columns); 4 epilogue warps read it with tcgen05.ld (lane =
//     field column), undo the scale, add h and write Gp with 128-byte coalesced stores.
// Warp roles: 0-3 epilogue, 4-11 one-hot producers, 12 A loader, 13 MMA issuer (highest warp id: top issue priority).
// Persistent grid, tiles ordered column-tile-major so concurrently running CTAs stream the same Jsym tiles from L2.
#include "common.cuh"
#include "../../include/ppde_b200.h"
#include "launch.cuh"
#include "tc_common.cuh"

namespace ppde {
namespace tc {

constexpr int PD_M = 128;                          // field columns per tile
constexpr int PD_N = 256;                          // chains per tile
constexpr int PD_STAGES = 3;
constexpr int PD_A_BYTES = 2 * PD_M * KCH * 2;     // hi + lo [128 x 64] fp16 = 32 KB
constexpr int PD_B_BYTES = PD_N * KCH * 2;         // [256 x 64] fp16 = 32 KB
constexpr int PD_STAGE_BYTES = PD_A_BYTES + PD_B_BYTES;
constexpr int PD_NTHREADS = 14 * 32;
constexpr int PD_WARP_LOAD = 12, PD_WARP_MMA = 13;

// D[tmem] (+)= A[smem desc] * B[smem desc]^T   (both K-major fp16, fp32 accumulate)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

// Jsym (fp32 [D, D]) -> tiled fp16 hi/lo image: block (mt, kc) = [hi 128x64 | lo 128x64], SW128 K-major.
__global__ void potts_dense_pack_kernel(const float* __restrict__ Jsym, int D, float jscale, int KC, int MT,
                                        unsigned char* __restrict__ Jt) {
    const int64_t total = (int64_t)MT * KC * PD_M * (KCH / 2);        // one thread per pair of k
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
        const int kp = (int)(o % (KCH / 2));
        const int r = (int)((o / (KCH / 2)) % PD_M);
        const int64_t blk = o / ((KCH / 2) * PD_M);                   // mt * KC + kc
        const int kc = (int)(blk % KC), mt = (int)(blk / KC);
        const int col = mt * PD_M + r, k0 = kc * KCH + 2 * kp;
        float v0 = 0.f, v1 = 0.f;
        if (col < D) {
            if (k0 < D) v0 = Jsym[(int64_t)col * D + k0] * jscale;
            if (k0 + 1 < D) v1 = Jsym[(int64_t)col * D + k0 + 1] * jscale;
        }
        const float h0 = h_trunc(v0), h1 = h_trunc(v1);
        const int k = 2 * kp;
        const int off = (r >> 3) * 1024 + (r & 7) * 128 + (((k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
        unsigned char* base = Jt + blk * PD_A_BYTES;
        *reinterpret_cast<uint32_t*>(base + off) = pack_h2(h0, h1);
        *reinterpret_cast<uint32_t*>(base + PD_M * KCH * 2 + off) = pack_h2(v0 - h0, v1 - h1);
    }
}

struct DenseParams {
    ppde_potts_t m;
    const unsigned char* Jt;
    float unscale;
    const uint8_t* aa;
    int aa_stride;
    int n;
    float* Gp;
    int64_t Gp_stride;
    int KC, MT, NTc;          // K chunks, column tiles, chain tiles
    int last_ksteps;          // K steps (of 16) in the last chunk
};

__global__ void __launch_bounds__(PD_NTHREADS, 1) potts_dense_tc_kernel(const __grid_constant__ DenseParams prm) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* stg = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(stg + PD_STAGES * PD_STAGE_BYTES);
    uint64_t* a_full = bars;                       // [PD_STAGES] bulk copy landed (tx bytes)
    uint64_t* b_full = a_full + PD_STAGES;         // [PD_STAGES] 8 producer warps
    uint64_t* empty = b_full + PD_STAGES;          // [PD_STAGES] tcgen05.commit
    uint64_t* dfull = empty + PD_STAGES;           // [2]
    uint64_t* dempty = dfull + 2;                  // [2] 4 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = prm.m.D, KC = prm.KC;
    const int total_tiles = prm.MT * prm.NTc;
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        for (int s = 0; s < PD_STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&b_full[s], 8); mbar_init(&empty[s], 1); }
        for (int d = 0; d < 2; ++d) { mbar_init(&dfull[d], 1); mbar_init(&dempty[d], 4); }
        fence_barrier_init();
    }
    if (warp == PD_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===== EPILOGUE: lane = field column; 32 chains per TMEM load; one 128-byte store per warp and chain =====
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int mt = tile / prm.NTc, nt = tile - mt * prm.NTc;
            const int buf = ti & 1;
            const int col = mt * PD_M + warp * 32 + lane;
            const float hv = (col < D) ? prm.m.h[col] : 0.f;
            mbar_wait(&dfull[buf], (uint32_t)((ti >> 1) & 1));
            tc_fence_after();
            const int chain0 = nt * PD_N;
#pragma unroll 1
            for (int cg = 0; cg < PD_N / 32; ++cg) {
                uint32_t r[32];
                tmem_ld32(lane_addr + buf * PD_N + cg * 32, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (col < D) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const int b = chain0 + cg * 32 + i;
                        if (b < prm.n) prm.Gp[(int64_t)b * prm.Gp_stride + col] = fmaf(__uint_as_float(r[i]), prm.unscale, hv);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&dempty[buf]);
        }
    } else if (warp == PD_WARP_MMA) {
        // ===== MMA ISSUER (warp-uniform loop, one elected lane issues) =====
        const uint32_t idesc = make_idesc(PD_M, PD_N);
        const uint32_t stg_addr = smem_u32(stg);
        int s = 0;
        uint32_t ph = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int buf = ti & 1;
            if (ti >= 2) mbar_wait(&dempty[buf], (uint32_t)(((ti >> 1) + 1) & 1));
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + buf * PD_N;
#pragma unroll 1
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(&a_full[s], ph);
                mbar_wait(&b_full[s], ph);
                tc_fence_after();
                const uint32_t sa = stg_addr + s * PD_STAGE_BYTES;
                const uint64_t ahi = make_b_desc(sa), alo = make_b_desc(sa + PD_M * KCH * 2), bd = make_b_desc(sa + PD_A_BYTES);
                const int ksteps = (kc == KC - 1) ? prm.last_ksteps : KCH / 16;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < KCH / 16; ++ks) {
                        if (ks < ksteps) {
                            mma_ss(d_tmem, ahi + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, (kc | ks) ? 1u : 0u);
                            mma_ss(d_tmem, alo + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), idesc, 1u);
                        }
                    }
                    tc_commit(&empty[s]);
                    if (kc == KC - 1) tc_commit(&dfull[buf]);
                }
                __syncwarp();
                if (++s == PD_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == PD_WARP_LOAD) {
        // ===== A LOADER: one 32 KB bulk copy per stage =====
        int s = 0;
        uint32_t ph = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int mt = tile / prm.NTc;
            const unsigned char* src = prm.Jt + (int64_t)mt * KC * PD_A_BYTES;
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(&empty[s], ph ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&a_full[s], PD_A_BYTES);
                    bulk_g2s(stg + s * PD_STAGE_BYTES, src + (int64_t)kc * PD_A_BYTES, PD_A_BYTES, &a_full[s]);
                }
                __syncwarp();
                if (++s == PD_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ===== ONE-HOT PRODUCERS: thread = chain row of the tile =====
        const int r = threadIdx.x - 4 * 32;                              // 0..255
        const uint32_t row_a = smem_u32(stg) + PD_A_BYTES + (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        const uint32_t sw = (uint32_t)(r & 7);
        int s = 0;
        uint32_t ph = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
            const int mt = tile / prm.NTc, nt = tile - mt * prm.NTc;
            const int b = nt * PD_N + r;
            const uint8_t* a = prm.aa + (int64_t)min(b, prm.n - 1) * prm.aa_stride + prm.m.win_lo;
            const bool live = b < prm.n;
            for (int kc = 0; kc < KC; ++kc) {
                // entries (i, a_i) with 20 i + a_i in [64 kc, 64 kc + 64): positions i0 .. i0 + 3
                const int k0 = kc * KCH;
                const int i0 = k0 / PPDE_Q;
                int rel[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int i = i0 + q;
                    rel[q] = (live && i < prm.m.Lp) ? i * PPDE_Q + (int)a[i] - k0 : -1;
                }
                mbar_wait(&empty[s], ph ^ 1);
                const uint32_t ra = row_a + (uint32_t)(s * PD_STAGE_BYTES);
                // zero the row: lane l starts at 16-byte unit l % 8, so the 8 rows of a store phase hit 8 different bank groups
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ra + ((uint32_t)((u + lane) & 7) << 4)), "r"(0u) : "memory");
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (rel[q] >= 0 && rel[q] < KCH) {
                        const uint32_t addr = ra + ((((uint32_t)rel[q] >> 3) ^ sw) << 4) + (((uint32_t)rel[q] & 7u) << 1);
                        asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)0x3C00) : "memory");   // 1.0h
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[s]);
                if (++s == PD_STAGES) { s = 0; ph ^= 1; }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == PD_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

// Epotts[b] = 1/2 sum_i ( Gp[b,(i,aa_i)] + h[(i,aa_i)] ) - wt_H     (one warp per chain)
__global__ void potts_energy_from_field_kernel(ppde_potts_t m, const uint8_t* __restrict__ aa, int aa_stride, int n,
                                               const float* __restrict__ Gp, int64_t Gp_stride,
                                               const int32_t* __restrict__ rows, float* __restrict__ Epotts) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n) return;
    const int lane = threadIdx.x & 31;
    const uint8_t* a = aa + (int64_t)b * aa_stride + m.win_lo;
    const float* g = Gp + (int64_t)(rows ? rows[b] : b) * Gp_stride;
    float part = 0.f;
    for (int i = lane; i < m.Lp; i += 32) {
        const int r = i * PPDE_Q + a[i];
        part += g[r] + m.h[r];
    }
    // fixed-order tree: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) Epotts[b] = 0.5f * part - m.wt_H;
}

// Oracle model (AugmentedLinearRegression, ppde/nets.py:315-347): mean over the ridge heads of
//   W_h[0] * sqrt(1/reg_potts) * dH(x) + sqrt(1/r_h) * sum_{(i,a)} W_h[1 + 20 i + a] x[i,a] + b_h .
// The heads are linear, so their mean is one head with averaged weights (folded on the host in fp64):
//   y[b] = sbar * dH[b] + sum_i wbar[20 i + aa_i] + cbar          (one warp per chain, fixed-order reduction)
__global__ void oracle_ridge_kernel(const float* __restrict__ wbar, float sbar, float cbar, const uint8_t* __restrict__ aa,
                                    int aa_stride, int n, int L, const float* __restrict__ dH, float* __restrict__ out) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= n) return;
    const int lane = threadIdx.x & 31;
    const uint8_t* a = aa + (int64_t)b * aa_stride;
    float part = 0.f;
    for (int i = lane; i < L; i += 32) part += __ldg(wbar + i * PPDE_Q + a[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) out[b] = fmaf(sbar, dH ? dH[b] : 0.f, part) + cbar;
}

}  // namespace tc
}  // namespace ppde

using namespace ppde;

extern "C" int ppde_potts_energy_rows(const ppde_potts_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n, const float* Gp,
                                      int64_t Gp_stride, const int32_t* rows, float* Epotts, void* stream) {
    if (n <= 0) return 0;
    tc::potts_energy_from_field_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(*m, aa, aa_stride, n, Gp, Gp_stride, rows, Epotts);
    return launch_done();
}

extern "C" int ppde_oracle_ridge(const float* wbar, float sbar, float cbar, const uint8_t* aa, int32_t aa_stride, int32_t n,
                                 int32_t L, const float* dH, float* out, void* stream) {
    if (n <= 0) return 0;
    tc::oracle_ridge_kernel<<<(n + 7) / 8, 256, 0, (cudaStream_t)stream>>>(wbar, sbar, cbar, aa, aa_stride, n, L, dH, out);
    return launch_done();
}

extern "C" int64_t ppde_potts_dense_image_bytes(int32_t D) {
    const int64_t KC = (D + tc::KCH - 1) / tc::KCH, MT = (D + tc::PD_M - 1) / tc::PD_M;
    return MT * KC * (int64_t)tc::PD_A_BYTES;
}

extern "C" int ppde_potts_dense_pack(const ppde_potts_t* m, float jscale, void* Jt, void* stream) {
    const int D = m->D;
    if (D <= 0 || !Jt) return (int)cudaErrorInvalidValue;
    const int KC = (D + tc::KCH - 1) / tc::KCH, MT = (D + tc::PD_M - 1) / tc::PD_M;
    tc::potts_dense_pack_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(m->Jsym, D, jscale, KC, MT, (unsigned char*)Jt);
    return launch_done();
}

extern "C" int ppde_potts_dense_full(const ppde_potts_t* m, const void* Jt, float jscale, const uint8_t* aa, int32_t aa_stride,
                                     int32_t n, float* Gp, int64_t Gp_stride, float* Epotts, void* stream) {
    if (n <= 0) return 0;
    if (m->D != m->Lp * PPDE_Q || !Jt || !(jscale > 0.f)) return (int)cudaErrorInvalidValue;
    tc::DenseParams prm;
    prm.m = *m;
    prm.Jt = (const unsigned char*)Jt;
    prm.unscale = 1.f / jscale;
    prm.aa = aa; prm.aa_stride = aa_stride; prm.n = n;
    prm.Gp = Gp; prm.Gp_stride = Gp_stride;
    prm.KC = (m->D + tc::KCH - 1) / tc::KCH;
    prm.MT = (m->D + tc::PD_M - 1) / tc::PD_M;
    prm.NTc = (n + tc::PD_N - 1) / tc::PD_N;
    const int klast = m->D - (prm.KC - 1) * tc::KCH;
    prm.last_ksteps = (klast + 15) / 16;
    const int sms = sm_count();
    const int tiles = prm.MT * prm.NTc;
    const int grid = tiles < sms ? tiles : sms;
    const size_t smem = 1024 + (size_t)tc::PD_STAGES * tc::PD_STAGE_BYTES + 32 * sizeof(uint64_t);
    static SmemCache configured;
    if (cudaError_t e = ensure_dynamic_smem(tc::potts_dense_tc_kernel, smem, configured)) return (int)e;
    cudaStream_t st = (cudaStream_t)stream;
    tc::potts_dense_tc_kernel<<<grid, tc::PD_NTHREADS, smem, st>>>(prm);
    int r = launch_done();
    if (r) return r;
    if (Epotts) {
        tc::potts_energy_from_field_kernel<<<(n + 7) / 8, 256, 0, st>>>(*m, aa, aa_stride, n, Gp, Gp_stride, nullptr, Epotts);
        r = launch_done();
    }
    return r;
}
