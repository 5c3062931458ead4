// Path-auxiliary proposal, reverse proposal and Metropolis-Hastings commit.
//
// Reference: PPDE_PAS.run inner loop             ppde/protein_samplers/ppde.py:65-153
//            mut_distance / mutation_mask          ppde/utils.py:5-28
//            safe_logits_to_probs                  ppde/utils.py:106-111
//            torch.distributions.Categorical       (__init__ renormalise, sample = multinomial =
//                                                   argmax p/Exp(1), log_prob = log(clamp(p)))
// Index-form semantics: SURVEY.md Appendix A.  One CTA per chain; the chain's gradient row
// [20L] lives in shared memory for all S sub-steps (it is frozen along the path, ppde.py:98).
#include "common.cuh"
#include "../../include/ppde_b200.h"
#include "launch.cuh"
#include <cstdlib>

namespace ppde {

__device__ __forceinline__ int y_row(int row_cur, int b, int n) { return row_cur == b ? n + b : b; }

// mutation_mask (ppde/utils.py:17-28) in index form: at or above the edit-distance threshold the only legal target at position
// i is "revert to the wild type", and only where the chain differs from it.  Returns the single allowed residue or -1.
__device__ __forceinline__ int revert_only_target(uint8_t cur, uint8_t wt) { return (cur != wt) ? (int)wt : -1; }

// softmax -> clamp -> renormalise statistics of one logit vector held in shared memory.
// On return sP[j] = clamp(exp(l_j - m1) / s2) and the function returns s3 = sum_j sP[j].
// (utils.py:106-111: logits - logsumexp, softmax, clamp_probs; Categorical.__init__: p / p.sum())
// `mx` is the caller's per-thread running maximum of the logits it wrote (saves one pass).
// exp through ex2.approx (2^-22 relative) and one reciprocal per vector instead of a division per entry: the
// probabilities agree with the reference's to ~2e-7 relative, far inside the 1e-4 tolerance; all reductions are
// fixed-order (deterministic).
template <int NT>
__device__ __forceinline__ float softmax_clamp_inplace(float* sP, int n4, float mx, float* red) {
    float4* p4 = reinterpret_cast<float4*>(sP);
    const float m1 = block_max<NT>(mx, red);
    const float off = (m1 == -INFINITY) ? 0.f : m1;
    float sum = 0.f;
    for (int q = threadIdx.x; q < n4; q += NT) {
        float4 v = p4[q];
        v.x = __expf(v.x - off); v.y = __expf(v.y - off); v.z = __expf(v.z - off); v.w = __expf(v.w - off);
        p4[q] = v;
        sum += (v.x + v.y) + (v.z + v.w);
    }
    const float s2 = block_sum<NT>(sum, red);
    const float r2 = 1.0f / s2;
    float sum3 = 0.f;
    for (int q = threadIdx.x; q < n4; q += NT) {
        float4 v = p4[q];
        v.x = clamp_prob(v.x * r2); v.y = clamp_prob(v.y * r2);
        v.z = clamp_prob(v.z * r2); v.w = clamp_prob(v.w * r2);
        p4[q] = v;
        sum3 += (v.x + v.y) + (v.z + v.w);
    }
    return block_sum<NT>(sum3, red);
}

template <int NT>
__global__ void __launch_bounds__(NT) pas_propose_kernel(ppde_potts_t m, ppde_chains_t c, ppde_pas_params_t p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = c.L, NE = L * PPDE_Q, n4 = NE / 4;
    float* sG = reinterpret_cast<float*>(smem_raw);          // [20L] gradient row (frozen)
    float* sP = sG + NE;                                     // [20L] logits -> probabilities
    float* sCur = sP + NE;                                   // [L]  G[i, aa_i]
    uint8_t* sAA = reinterpret_cast<uint8_t*>(sCur + L);     // [L]  evolving state
    uint8_t* sWT = sAA + ((L + 15) & ~15);                   // [L]
    __shared__ float red[33];
    __shared__ int redi[33];
    __shared__ int s_idx;
    __shared__ int s_best;                                   // bits of the best race quotient so far (positive floats order as ints)

    const int b = blockIdx.x;
    const uint32_t gid = (uint32_t)(c.chain_offset + b);
    const int t = p.t_dev ? *p.t_dev : p.t;
    const Philox rng(p.seed);
    const int lo = p.min_pos, hi = p.max_pos;

    {   // stage the row and the state
        const float4* g4 = reinterpret_cast<const float4*>(c.G + (int64_t)c.row_cur[b] * NE);
        float4* s4 = reinterpret_cast<float4*>(sG);
        for (int q = threadIdx.x; q < n4; q += NT) s4[q] = g4[q];
        const uint8_t* ax = c.aa + (int64_t)b * c.aa_stride;
        for (int i = threadIdx.x; i < L; i += NT) { sAA[i] = ax[i]; sWT[i] = m.wt[i]; }
    }
    const int span = p.S;                                      // 2*pas-1 values: U in {1..S}  (ppde.py:67)
    const int U = 1 + (int)(rng(0u, gid, (uint32_t)t, (uint32_t)(KIND_PATHLEN << 16)).x % (uint32_t)span);
    if (threadIdx.x == 0) c.U[b] = U;
    __syncthreads();
    // sub-steps s >= U are masked out by the reference (u_mask, ppde.py:111-115,132): they are evaluated only on request
    const int S_eff = p.full_trace ? p.S : U;
    for (int s = S_eff + (int)threadIdx.x; s < p.S; s += NT) {
        const int64_t o = (int64_t)s * c.n + b;
        c.idx[o] = -1; c.old_aa[o] = 0; c.lqf[o] = 0.f;
    }

    for (int s = 0; s < S_eff; ++s) {
        // edit distance to WT and the threshold flag (utils.py:5-14, ppde.py:86-91)
        int dpart = 0;
        for (int i = threadIdx.x; i < L; i += NT) {
            dpart += (sAA[i] != sWT[i]);
            sCur[i] = sG[i * PPDE_Q + sAA[i]];
        }
        const int dist = block_sum_int<NT>(dpart, redi);       // (barrier inside also publishes sCur)
        const bool at_thr = dist >= p.nmut_threshold;
        // Taylor logits with the revert-only mask and the window mask (ppde.py:95-104, utils.py:17-28)
        float4* p4 = reinterpret_cast<float4*>(sP);
        const float4* g4 = reinterpret_cast<const float4*>(sG);
        float lmax = -INFINITY;
        for (int q = threadIdx.x; q < n4; q += NT) {
            const int i = q / 5, a0 = (q - i * 5) * 4;
            float4 g = g4[q];
            const float gc = sCur[i];
            float4 l;
            l.x = (g.x - gc) * 0.5f; l.y = (g.y - gc) * 0.5f; l.z = (g.z - gc) * 0.5f; l.w = (g.w - gc) * 0.5f;
            if (i < lo || i > hi) {
                l.x = l.y = l.z = l.w = -INFINITY;
            } else if (at_thr) {
                const int w = revert_only_target(sAA[i], sWT[i]);      // the only legal target: revert to WT
                if (a0 + 0 != w) l.x = -INFINITY;
                if (a0 + 1 != w) l.y = -INFINITY;
                if (a0 + 2 != w) l.z = -INFINITY;
                if (a0 + 3 != w) l.w = -INFINITY;
            }
            p4[q] = l;
            lmax = fmaxf(fmaxf(lmax, fmaxf(l.x, l.y)), fmaxf(l.z, l.w));
        }
        if (threadIdx.x == 0) s_best = 0;
        const float s3 = softmax_clamp_inplace<NT>(sP, n4, lmax, red);     // (its barriers also publish s_best = 0)
        // exponential race: argmax_j p_j / E_j, E_j = -log(u_j)  (= torch.multinomial(p, 1, True)).
        // The quotient r_j = (p_j / s3) / (-log u_j) is evaluated exactly as before, but only for entries that can
        // still win: -log u >= 1 - u, so r_j <= p_j / (s3 (1 - u_j)); an entry whose bound is below the best quotient
        // seen so far by ANY thread of the block (s_best, monotone) with a 1e-5 margin (>> rounding) is skipped.
        float best = -1.f; int bidx = 0x7fffffff;
        const float* um = p.uniforms ? p.uniforms + ((int64_t)s * c.n + b) * NE : nullptr;
        for (int q = threadIdx.x; q < n4; q += NT) {
            // the cheap test first, on the raw Philox words: entry j can still win only if p_j (1 + 1e-5) > thr (1 - u_j), and
            // 1 - u_j = ((~x_j >> 8) + 0.5) 2^-24 needs no uniform; one branch per float4, taken for a handful of entries per vector
            const float4 pr = p4[q];
            const float thr = fmaxf(best, __int_as_float(*(volatile int*)&s_best)) * s3 * (1.0f / 1.00001f);
            uint4 w = make_uint4(0u, 0u, 0u, 0u);
            float4 u1;                                                       // 1 - u
            if (um) {
                const float4 u = reinterpret_cast<const float4*>(um)[q];
                u1 = make_float4(1.0f - u.x, 1.0f - u.y, 1.0f - u.z, 1.0f - u.w);
            } else {
                w = rng((uint32_t)q, gid, (uint32_t)t, (uint32_t)(s | (KIND_PROPOSAL << 16)));
                u1 = make_float4(u32_to_unit(~w.x), u32_to_unit(~w.y), u32_to_unit(~w.z), u32_to_unit(~w.w));
            }
            const bool c0 = pr.x > thr * u1.x, c1 = pr.y > thr * u1.y, c2 = pr.z > thr * u1.z, c3 = pr.w > thr * u1.w;
            if (c0 | c1 | c2 | c3) {                                         // may still win: exact evaluation, in entry order
                float4 u;
                if (um) u = reinterpret_cast<const float4*>(um)[q];
                else u = make_float4(u32_to_unit(w.x), u32_to_unit(w.y), u32_to_unit(w.z), u32_to_unit(w.w));
                const float pv[4] = {pr.x, pr.y, pr.z, pr.w};
                const float uv[4] = {u.x, u.y, u.z, u.w};
                const bool cv[4] = {c0, c1, c2, c3};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (cv[k]) {
                        const float r = (pv[k] / s3) / (-logf(uv[k]));
                        if (r > best) { best = r; bidx = q * 4 + k; atomicMax(&s_best, __float_as_int(r)); }
                    }
                }
            }
        }
        block_argmax<NT>(best, bidx, red, redi);
        if (threadIdx.x == 0) {
            const int pos = bidx / PPDE_Q, a = bidx - pos * PPDE_Q;
            const int64_t o = (int64_t)s * c.n + b;
            c.idx[o] = bidx;
            c.old_aa[o] = sAA[pos];
            c.lqf[o] = logf(clamp_prob(sP[bidx] / s3));            // Categorical.log_prob
            if (s < U) sAA[pos] = (uint8_t)a;                      // ppde.py:111-115 (masked by u_mask)
            s_idx = bidx;
        }
        __syncthreads();
    }
    uint8_t* ay = c.aa_y + (int64_t)b * c.aa_stride;
    for (int i = threadIdx.x; i < L; i += NT) ay[i] = sAA[i];
}

template <int NT>
__global__ void __launch_bounds__(NT) pas_reverse_accept_kernel(ppde_potts_t m, ppde_chains_t c,
                                                                ppde_pas_params_t p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int L = c.L, NE = L * PPDE_Q, n4 = NE / 4;
    float* sG = reinterpret_cast<float*>(smem_raw);          // [20L] gradient row at y
    float* sP = sG + NE;                                     // [20L]
    float* sCur = sP + NE;                                   // [L]
    uint8_t* sZ = reinterpret_cast<uint8_t*>(sCur + L);      // [L] trajectory state
    __shared__ float red[33];
    __shared__ int redi[33];
    __shared__ int s_flag;

    const int b = blockIdx.x, n = c.n;
    const uint32_t gid = (uint32_t)(c.chain_offset + b);
    const int t = p.t_dev ? *p.t_dev : p.t;
    const int rx = c.row_cur[b];
    const int ry = y_row(rx, b, n);
    uint8_t* ax = c.aa + (int64_t)b * c.aa_stride;
    const uint8_t* ay = c.aa_y + (int64_t)b * c.aa_stride;
    {
        const float4* g4 = reinterpret_cast<const float4*>(c.G + (int64_t)ry * NE);
        float4* s4 = reinterpret_cast<float4*>(sG);
        for (int q = threadIdx.x; q < n4; q += NT) s4[q] = g4[q];
        for (int i = threadIdx.x; i < L; i += NT) sZ[i] = ax[i];
    }
    const int U = c.U[b];
    __syncthreads();

    float log_ratio = 0.f;
    const int S_eff = p.full_trace ? p.S : U;                      // dead sub-steps (s >= U) only on request
    for (int s = S_eff + (int)threadIdx.x; s < p.S; s += NT) c.lqr[(int64_t)s * n + b] = 0.f;
    for (int s = 0; s < S_eff; ++s) {
        const int64_t o = (int64_t)s * n + b;
        const int cidx = c.idx[o];
        if (threadIdx.x == 0 && s < U) sZ[cidx / PPDE_Q] = (uint8_t)(cidx % PPDE_Q);   // state AFTER move s
        __syncthreads();
        for (int i = threadIdx.x; i < L; i += NT) sCur[i] = sG[i * PPDE_Q + sZ[i]];
        __syncthreads();
        float4* p4 = reinterpret_cast<float4*>(sP);
        const float4* g4 = reinterpret_cast<const float4*>(sG);
        float lmax = -INFINITY;
        for (int q = threadIdx.x; q < n4; q += NT) {              // NO masks on the reverse path (ppde.py:126-127)
            const float gc = sCur[q / 5];
            float4 g = g4[q];
            const float4 l = make_float4((g.x - gc) * 0.5f, (g.y - gc) * 0.5f, (g.z - gc) * 0.5f, (g.w - gc) * 0.5f);
            p4[q] = l;
            lmax = fmaxf(fmaxf(lmax, fmaxf(l.x, l.y)), fmaxf(l.z, l.w));
        }
        const float s3 = softmax_clamp_inplace<NT>(sP, n4, lmax, red);
        if (threadIdx.x == 0) {
            const float lqr = logf(clamp_prob(sP[cidx] / s3));
            c.lqr[o] = lqr;
            if (s < U) log_ratio += lqr - c.lqf[o];               // u_mask * (rev - fwd), ppde.py:132
        }
        __syncthreads();
    }

    if (threadIdx.x == 0) {
        const float e_x = c.E[b], f_x = c.fit[b];
        const float e_y = c.E_y[b], f_y = c.fit_y[b];
        const float log_acc = (e_y - e_x) + log_ratio;             // ppde.py:135-136
        const float u = u32_to_unit(Philox(p.seed)(0u, gid, (uint32_t)t, (uint32_t)(KIND_ACCEPT << 16)).x);
        const bool acc = expf(log_acc) >= u;                       // '>=' (ppde.py:138); NaN rejects
        c.log_acc[b] = log_acc;
        c.accept[b] = acc ? 1 : 0;
        const float e_rec = acc ? e_y : e_x, f_rec = acc ? f_y : f_x;   // ppde.py:141-143
        if (c.E_hist) c.E_hist[(int64_t)(t + 1) * n + b] = e_rec;
        if (c.fit_hist) c.fit_hist[(int64_t)(t + 1) * n + b] = f_rec;
        // 0: keep current state; 1: take y; 2: fall back to the paper-mode anchor
        s_flag = acc ? 1 : (p.paper_results ? 2 : 0);
        if (acc) { c.E[b] = e_y; c.fit[b] = f_y; c.row_cur[b] = ry; }
        else if (p.paper_results) {                                // x is never refreshed: reject = back to x0 (ppde.py:76-77,139)
            const int f = c.anchor_fixed ? c.anchor_fixed[b] : (c.row_wt - 2 * n);
            c.E[b] = c.E_fixed[f]; c.fit[b] = c.fit_fixed[f]; c.row_cur[b] = 2 * n + f;
        }
        const bool better = e_rec > c.best_E[b];                   // strict: first occurrence of the max (ppde.py:173)
        if (better) { c.best_E[b] = e_rec; c.best_fit[b] = f_rec; }
        s_flag |= better ? 4 : 0;
    }
    __syncthreads();
    const int flag = s_flag & 3;
    const bool better = (s_flag & 4) != 0;
    const uint8_t* src = ax;
    if (flag == 1) src = ay;
    else if (flag == 2) {
        const int f = c.anchor_fixed ? c.anchor_fixed[b] : (c.row_wt - 2 * n);
        src = c.aa_fixed + (int64_t)f * c.aa_stride;
    }
    // recorded state (post-accept, pre-reset): best-of-history and the random trajectory (ppde.py:142,146,172-183)
    int dpart = 0;
    for (int i = threadIdx.x; i < L; i += NT) {
        const uint8_t v = src[i];
        sZ[i] = v;
        dpart += (v != m.wt[i]);
        if (better) c.best_aa[(int64_t)b * c.aa_stride + i] = v;
        if (c.traj_aa && b == c.traj_chain) c.traj_aa[(int64_t)(t + 1) * c.aa_stride + i] = v;
    }
    const int dist = block_sum_int<NT>(dpart, redi);
    const bool reset = !p.paper_results && dist >= p.nmut_threshold;   // hard reset to WT (ppde.py:148-153)
    for (int i = threadIdx.x; i < L; i += NT) ax[i] = reset ? m.wt[i] : sZ[i];
    if (reset && threadIdx.x == 0) {
        const int f = c.row_wt - 2 * n;
        c.E[b] = c.E_fixed[f]; c.fit[b] = c.fit_fixed[f]; c.row_cur[b] = c.row_wt;
    }
}

// ======================================================================================================================
// Block reductions with ONE barrier each (every reduction has its own scratch), used by the position-per-thread kernels below.
// Round 1 had strided register-resident variants of the two kernels here (float4 index q = tid + k * 256 per thread); the
// position-per-thread kernels replaced them.
constexpr int PAS_NT = 256;

template <int NW> __device__ __forceinline__ float red_max1(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) r = fmaxf(r, red[w]);
    return r;
}
template <int NW> __device__ __forceinline__ float red_sum1(float v, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) r += red[w];
    return r;
}
template <int NW> __device__ __forceinline__ int red_sum1i(int v, int* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    int r = red[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) r += red[w];
    return r;
}
template <int NW> __device__ __forceinline__ void red_argmax1(float& v, int& idx, float* redv, int* redi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = v; redi[threadIdx.x >> 5] = idx; }
    __syncthreads();
    v = redv[0]; idx = redi[0];
#pragma unroll
    for (int w = 1; w < NW; ++w) {
        const float ov = redv[w]; const int oi = redi[w];
        if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
}

// ======================================================================================================================
// Position-per-thread variants (L <= 256; default).  Thread i of a 256-thread CTA owns position i: its 20 gradient entries,
// logits and probabilities live in REGISTERS for the whole kernel and are indexed by compile-time constants only.
//   * no shared-memory copy of the row, no `q < n4` predicates, no q / 5 index arithmetic, G[i, cur_i] and the state of the
//     position are thread-local (a move updates ONE thread's registers: no barrier for it);
//   * the logits are fma(g, 0.5, -0.5 gc) - bit-identical to (g - gc) * 0.5: scaling by 0.5 is exact;
//   * Philox counters are the same float4 indices q = 5 i + k as in the kernels above: same streams, entry for entry;
//   * 3 barriers per reverse sub-step, 5 per forward sub-step; block sums are fixed-order (deterministic).
// ncu (r02, fused strided kernel): 20 k warp-instructions per chain and 47 % of the warp time waiting for global loads;
// this layout executes about a third of that.
constexpr int PAS_CN = 3;                  // nets the fused gradient combine handles
constexpr int PAS_TRAIL = 16, PAS_MC32 = 64 * 32;     // tc::BD_TRAIL, tc::BD_MC * 32 (csrc/cnn_tc.cu): record trailer position

__device__ __forceinline__ float pick20(const float (&v)[PPDE_Q], int a) {
    float r = v[0];
#pragma unroll
    for (int k = 1; k < PPDE_Q; ++k) r = (a == k) ? v[k] : r;
    return r;
}
__device__ __forceinline__ void load_row20(float (&g)[PPDE_Q], const float* src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const float4 v = __ldcs(s4 + k);
        g[4 * k] = v.x; g[4 * k + 1] = v.y; g[4 * k + 2] = v.z; g[4 * k + 3] = v.w;
    }
}
// softmax -> clamp -> renormalise of the logits l (thread-local, -inf for masked / inactive entries): on return
// e[k] = clamp(exp(l - m1) / s2) for an active thread; returns s3 = sum over the block.  Same formulas as softmax_clamp_inplace.
template <int NW>
__device__ __forceinline__ float softmax_clamp_pos(float (&e)[PPDE_Q], bool active, float* red) {
    float lmax = -INFINITY;
#pragma unroll
    for (int k = 0; k < PPDE_Q; ++k) lmax = fmaxf(lmax, e[k]);
    const float m1 = red_max1<NW>(lmax, red);
    const float off = (m1 == -INFINITY) ? 0.f : m1;
    float sum = 0.f;
    if (active) {
#pragma unroll
        for (int k = 0; k < PPDE_Q; k += 4) {
            e[k] = __expf(e[k] - off); e[k + 1] = __expf(e[k + 1] - off); e[k + 2] = __expf(e[k + 2] - off); e[k + 3] = __expf(e[k + 3] - off);
            sum += (e[k] + e[k + 1]) + (e[k + 2] + e[k + 3]);
        }
    }
    const float s2 = red_sum1<NW>(sum, red + NW);
    const float r2 = 1.0f / s2;
    float sum3 = 0.f;
    if (active) {
#pragma unroll
        for (int k = 0; k < PPDE_Q; k += 4) {
            e[k] = clamp_prob(e[k] * r2); e[k + 1] = clamp_prob(e[k + 1] * r2);
            e[k + 2] = clamp_prob(e[k + 2] * r2); e[k + 3] = clamp_prob(e[k + 3] * r2);
            sum3 += (e[k] + e[k + 1]) + (e[k + 2] + e[k + 3]);
        }
    }
    return red_sum1<NW>(sum3, red + 2 * NW);
}

// The same statistics for UNMASKED Taylor logits l_k = fma(g_k, 0.5, hgc) of a thread's position (hgc = -0.5 G[i, cur_i], or -inf
// for a position outside the window / an inactive thread), straight from the gradient entries:
//   * the thread's largest logit is fma(gmax, 0.5, hgc) with gmax = max_k g_k, constant along the path (fma is monotone in g);
//   * exp(l_k - m1) = ex2(fma(g_k, 0.5 log2e, (hgc - m1) log2e)): one FFMA + one MUFU per entry, the logits are never stored
//     (the argument is rounded once instead of three times; ex2.approx itself is 2^-22 relative).
template <int NW>
__device__ __forceinline__ float softmax_clamp_pos_fast(float (&e)[PPDE_Q], const float (&g)[PPDE_Q], float hgc, float gmax,
                                                        bool active, float* red) {
    const float m1 = red_max1<NW>(fmaf(gmax, 0.5f, hgc), red);
    const float off = (m1 == -INFINITY) ? 0.f : m1;
    constexpr float LOG2E = 1.4426950408889634f;
    const float c2 = (hgc - off) * LOG2E;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < PPDE_Q; k += 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[k + j]) : "f"(fmaf(g[k + j], 0.5f * LOG2E, c2)));
        sum += (e[k] + e[k + 1]) + (e[k + 2] + e[k + 3]);
    }
    const float s2 = red_sum1<NW>(sum, red + NW);
    const float r2 = 1.0f / s2;
    float sum3 = 0.f;
    if (active) {
#pragma unroll
        for (int k = 0; k < PPDE_Q; k += 4) {
            e[k] = clamp_prob(e[k] * r2); e[k + 1] = clamp_prob(e[k + 1] * r2);
            e[k + 2] = clamp_prob(e[k + 2] * r2); e[k + 3] = clamp_prob(e[k + 3] * r2);
            sum3 += (e[k] + e[k + 1]) + (e[k + 2] + e[k + 3]);
        }
    }
    return red_sum1<NW>(sum3, red + 2 * NW);
}

// UM: the proposal uniforms come from a caller-provided array (teacher-forced parity tests) instead of the Philox streams
template <int MINB, bool UM>
__global__ void __launch_bounds__(PAS_NT, MINB) pas_propose_pos_kernel(ppde_potts_t m, ppde_chains_t c, ppde_pas_params_t p,
                                                                    const __grid_constant__ PhiloxKeys keys) {
    constexpr int NW = PAS_NT / 32;
    __shared__ float red[4 * NW];
    __shared__ int redi[2 * NW];
    __shared__ int s_best;                                   // bits of the best race quotient so far (positive floats order as ints)
    __shared__ int s_dist;
    const int L = c.L, NE = L * PPDE_Q;
    const int b = blockIdx.x, tid = threadIdx.x, i = tid;
    const bool active = i < L;
    const uint32_t gid = (uint32_t)(c.chain_offset + b);
    const int t = p.t_dev ? *p.t_dev : p.t;
    const Philox rng(p.seed);
    float g[PPDE_Q];
#pragma unroll
    for (int k = 0; k < PPDE_Q; ++k) g[k] = 0.f;
    int ai = 0, wi = 0;                                      // residue of my position: evolving state, wild type
    const int rx = c.row_cur[b];
    if (active) {
        load_row20(g, c.G + (int64_t)rx * NE + i * PPDE_Q);
        ai = c.aa[(int64_t)b * c.aa_stride + i];
        wi = m.wt[i];
    }
    const int ai0 = ai;
    __shared__ int s_mpos[PPDE_MAX_S];                       // positions of the applied moves, in sub-step order (fused Potts update)
    const int span = p.S;                                      // 2*pas-1 values: U in {1..S}  (ppde.py:67)
    const int U = 1 + (int)(rng(0u, gid, (uint32_t)t, (uint32_t)(KIND_PATHLEN << 16)).x % (uint32_t)span);
    if (tid == 0) c.U[b] = U;
    const int S_eff = p.full_trace ? p.S : U;                  // dead sub-steps (s >= U) only on request
    for (int s = S_eff + tid; s < p.S; s += PAS_NT) {
        const int64_t o = (int64_t)s * c.n + b;
        c.idx[o] = -1; c.old_aa[o] = 0; c.lqf[o] = 0.f;
    }
    {   // edit distance to WT (utils.py:5-14), once; updated per move below
        const int d0 = red_sum1i<NW>((active && ai != wi) ? 1 : 0, redi);
        if (tid == 0) s_dist = d0;
    }
    float hgc = -0.5f * pick20(g, ai);                         // -0.5 G[i, cur_i]
    const bool inwin = active && i >= p.min_pos && i <= p.max_pos;
    float gmax = g[0];
#pragma unroll
    for (int k = 1; k < PPDE_Q; ++k) gmax = fmaxf(gmax, g[k]);
    __syncthreads();

    for (int s = 0; s < S_eff; ++s) {
        const bool at_thr = s_dist >= p.nmut_threshold;       // ppde.py:86-91
        if (tid == 0) s_best = 0;
        // Taylor logits with the revert-only mask and the window mask (ppde.py:95-104, utils.py:17-28)
        float e[PPDE_Q];
        float s3;                                              // (the barriers of the softmax also publish s_best = 0)
        if (!at_thr) {                                         // only the window mask: a whole position is in or out
            s3 = softmax_clamp_pos_fast<NW>(e, g, inwin ? hgc : -INFINITY, gmax, active, red);
        } else {
            const int w = revert_only_target((uint8_t)ai, (uint8_t)wi);      // the only legal target: revert to WT
#pragma unroll
            for (int k = 0; k < PPDE_Q; ++k) e[k] = (inwin && w == k) ? fmaf(g[k], 0.5f, hgc) : -INFINITY;
            s3 = softmax_clamp_pos<NW>(e, active, red);
        }
        // exponential race: argmax_j p_j / E_j, E_j = -log(u_j)  (= torch.multinomial(p, 1, True)); an entry whose upper bound
        // p_j / (s3 (1 - u_j)) is below the best quotient seen so far by any thread (s_best, monotone), with a 1e-5 margin, cannot
        // win and is not evaluated; the quotient itself is evaluated exactly as in pas_propose_kernel.
        float best = -1.f; int bidx = 0x7fffffff;
        const uint32_t amask = __ballot_sync(0xffffffffu, active);
        if (active) {
            const float* um = UM ? p.uniforms + ((int64_t)s * c.n + b) * NE + i * PPDE_Q : nullptr;
            const uint32_t ctr3 = (uint32_t)(s | (KIND_PROPOSAL << 16));
            const float rs3 = 1.0f / s3;
            // Seed of the threshold: -log u <= (1 - u) / u, so r_j >= (p_j / s3) u_j / (1 - u_j); the largest such LOWER bound over
            // the first four entries of every thread (cheap: no logarithm, approximate reciprocal, 1e-5 margin) is published
            // before anything is evaluated exactly - otherwise every thread evaluates its first four entries (a fifth of the
            // vector) with two divisions and a logarithm each.  u >= k 2^-23 = f - 1 and 1 - u <= 2 - f, f = [1.k], k = x >> 9.
            uint4 wd0 = make_uint4(0u, 0u, 0u, 0u);
            if (!UM) {
                wd0 = philox_keyed(keys, (uint32_t)(5 * i), gid, (uint32_t)t, ctr3);
                const uint32_t ww[4] = {wd0.x, wd0.y, wd0.z, wd0.w};
                float lb = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float f = __uint_as_float((ww[j] >> 9) | 0x3F800000u);
                    float rc;
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(2.0f - f));
                    lb = fmaxf(lb, e[j] * (f - 1.0f) * rc);
                }
                lb *= rs3 * 0.99999f;
                const int lbw = __reduce_max_sync(amask, __float_as_int(lb));        // (non-negative floats order as ints)
                if ((tid & 31) == 0) atomicMax(&s_best, lbw);
            }
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float thr = fmaxf(best, __int_as_float(*(volatile int*)&s_best)) * s3 * (1.0f / 1.00001f);
                uint4 wd = wd0;
                float4 u1;                                                       // a lower bound of 1 - u (the test must never skip a winner)
                if (UM) {
                    const float4 u = reinterpret_cast<const float4*>(um)[k];
                    u1 = make_float4(1.0f - u.x, 1.0f - u.y, 1.0f - u.z, 1.0f - u.w);
                } else {
                    if (k) wd = philox_keyed(keys, (uint32_t)(5 * i + k), gid, (uint32_t)t, ctr3);
                    // 1 - u = 1 - ((x >> 8) + 0.5) 2^-24 > 1 - (k + 1) 2^-23 = (2 - 2^-23) - [1.k], k = x >> 9: exact subtraction
                    constexpr float C2 = 1.99999988079071044921875f;
                    u1 = make_float4(C2 - __uint_as_float((wd.x >> 9) | 0x3F800000u), C2 - __uint_as_float((wd.y >> 9) | 0x3F800000u),
                                     C2 - __uint_as_float((wd.z >> 9) | 0x3F800000u), C2 - __uint_as_float((wd.w >> 9) | 0x3F800000u));
                }
                const float t1[4] = {thr * u1.x, thr * u1.y, thr * u1.z, thr * u1.w};
                if ((e[4 * k] > t1[0]) | (e[4 * k + 1] > t1[1]) | (e[4 * k + 2] > t1[2]) | (e[4 * k + 3] > t1[3])) {   // may still win: exact evaluation, in entry order
                    asm volatile("" : "+r"(wd.x), "+r"(wd.y), "+r"(wd.z), "+r"(wd.w));   // (keeps the conversions below inside the branch)
                    float4 u;
                    if (UM) u = reinterpret_cast<const float4*>(um)[k];
                    else u = make_float4(u32_to_unit(wd.x), u32_to_unit(wd.y), u32_to_unit(wd.z), u32_to_unit(wd.w));
                    const float pq[4] = {e[4 * k], e[4 * k + 1], e[4 * k + 2], e[4 * k + 3]};
                    const float uq[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (pq[j] > t1[j]) {                                     // (the per-entry tests again: only here, in the rare branch)
                            const float r = (pq[j] / s3) / (-logf(uq[j]));
                            if (r > best) { best = r; bidx = i * PPDE_Q + 4 * k + j; atomicMax(&s_best, __float_as_int(r)); }
                        }
                    }
                }
            }
        }
        red_argmax1<NW>(best, bidx, red + 3 * NW, redi + NW);
        const int pos = bidx / PPDE_Q, a = bidx - pos * PPDE_Q;
        if (i == pos) {                                        // the thread that owns the position records the move
            const int64_t o = (int64_t)s * c.n + b;
            c.idx[o] = bidx;
            c.old_aa[o] = (uint8_t)ai;
            c.lqf[o] = logf(clamp_prob(pick20(e, a) / s3));                  // Categorical.log_prob
            if (s < U) {                                                     // ppde.py:111-115 (masked by u_mask)
                s_dist += (int)(a != wi) - (int)(ai != wi);
                ai = a;
                hgc = -0.5f * pick20(g, a);
                s_mpos[s] = pos;
            }
        }
        __syncthreads();                                       // s_dist, s_best of the next sub-step
    }
    if (active) c.aa_y[(int64_t)b * c.aa_stride + i] = (uint8_t)ai;
    if (p.fuse_potts) {
        // Potts field of the proposal (potts_incremental_kernel, same sums in the same order): Gp_y = Gp_x + sum over the net
        // changes of (Jsym[new] - Jsym[old]); thread = position, 5 x 16-byte loads per coupling row; Epotts_y from the new field.
        __shared__ int s_new[PPDE_MAX_S], s_old[PPDE_MAX_S];
        __shared__ int s_cnt;
        __shared__ uint8_t s_ax[PAS_NT], s_ay[PAS_NT];
        s_ax[tid] = (uint8_t)ai0; s_ay[tid] = (uint8_t)ai;
        __syncthreads();
        if (tid == 0) {
            int cnt = 0;
            for (int s = 0; s < p.S && s < U; ++s) {
                const int pos = s_mpos[s];
                if (pos < m.win_lo || pos >= m.win_lo + m.Lp) continue;       // outside the Potts window: no coupling
                if (s_ax[pos] == s_ay[pos]) continue;                         // no net change at this position
                const int rn = (pos - m.win_lo) * PPDE_Q + s_ay[pos];
                bool dup = false;
                for (int q = 0; q < cnt; ++q) dup |= (s_new[q] == rn);
                if (dup) continue;
                s_new[cnt] = rn;
                s_old[cnt] = (pos - m.win_lo) * PPDE_Q + s_ax[pos];
                ++cnt;
            }
            s_cnt = cnt;
        }
        __syncthreads();
        const int cnt = s_cnt;
        const int ry = y_row(rx, b, c.n);
        const int D4 = m.D / 4;
        const float4* J4 = reinterpret_cast<const float4*>(m.Jsym);
        const float4* gx = reinterpret_cast<const float4*>(c.Gp + (int64_t)rx * m.D);
        float4* gy = reinterpret_cast<float4*>(c.Gp + (int64_t)ry * m.D);
        for (int q = tid; q < D4; q += PAS_NT) {               // coalesced: consecutive threads, consecutive 16-byte words
            float4 a4 = __ldcs(gx + q);
            for (int k = 0; k < cnt; ++k) {
                const float4 vn = __ldg(J4 + (int64_t)s_new[k] * D4 + q);
                const float4 vo = __ldg(J4 + (int64_t)s_old[k] * D4 + q);
                a4.x += vn.x - vo.x; a4.y += vn.y - vo.y; a4.z += vn.z - vo.z; a4.w += vn.w - vo.w;
            }
            gy[q] = a4;
        }
        __syncthreads();
        const int ip = i - m.win_lo;
        float part = 0.f;
        if (active && ip >= 0 && ip < m.Lp) {
            const int r = ip * PPDE_Q + ai;
            part = c.Gp[(int64_t)ry * m.D + r] + m.h[r];
        }
        const float tot = red_sum1<NW>(part, red);
        if (tid == 0) c.Epotts_y[b] = 0.5f * tot - m.wt_H;
    }
}

__global__ void __launch_bounds__(PAS_NT, 4) pas_reverse_accept_pos_kernel(ppde_potts_t m, ppde_chains_t c, ppde_pas_params_t p) {
    constexpr int NT = PAS_NT, NW = PAS_NT / 32;
    __shared__ float red[3 * NW];
    __shared__ int redi[NW];
    __shared__ int s_flag;
    __shared__ float s_ratio;
    __shared__ uint32_t s_map[2][PAS_NT];                    // output row -> entry of the net's row list, tagged with the epoch
    __shared__ uint32_t s_map0[PAS_CN][PAS_NT];              // the same for the first tile of every net (entry + 1), built at once

    const int L = c.L, NE = L * PPDE_Q;
    const int b = blockIdx.x, n = c.n, tid = threadIdx.x, i = tid;
    const bool active = i < L;
    const uint32_t gid = (uint32_t)(c.chain_offset + b);
    const int t = p.t_dev ? *p.t_dev : p.t;
    const int rx = c.row_cur[b];
    const int ry = y_row(rx, b, n);
    uint8_t* ax = c.aa + (int64_t)b * c.aa_stride;
    const uint8_t* ay = c.aa_y + (int64_t)b * c.aa_stride;
    const int U = c.U[b];
    const int S_eff = p.full_trace ? p.S : U;                      // dead sub-steps (s >= U) only on request
    int cpre[4];                                                   // proposal indices of the first sub-steps (loaded early)
#pragma unroll
    for (int s = 0; s < 4; ++s) cpre[s] = (s < S_eff) ? c.idx[(int64_t)s * n + b] : 0;
    float g[PPDE_Q];
#pragma unroll
    for (int k = 0; k < PPDE_Q; ++k) g[k] = 0.f;
    int zi = active ? (int)ax[i] : 0;                              // trajectory state of my position
    if (tid == 0) s_ratio = 0.f;
    // what the accept / commit tail needs, requested now (three more dependent DRAM round trips at the end of the kernel otherwise)
    const int xi0 = zi;
    const int yi0 = active ? (int)ay[i] : 0;
    const int wti = active ? (int)m.wt[i] : 0;
    float pe_x = 0.f, pf_x = 0.f, pe_y = 0.f, pf_y = 0.f, pbest = 0.f;
    if (tid == 0) { pe_x = c.E[b]; pf_x = c.fit[b]; pe_y = c.E_y[b]; pf_y = c.fit_y[b]; pbest = c.best_E[b]; }
    if (p.comb_nets > 0) {
        // Fused gradient combine (delta backward): G_y = G_x + (Gp_y - Gp_x)(window), then for k = 0 .. nets-1, tile by tile:
        // G_y[row] += scale * dGc_k[row] for the output rows the record lists - the sparse changes cnn_backward_delta_kernel left
        // in the scratch.  Fixed order (net, tile): the same sums as cnn_grad_combine_sparse_kernel, bit for bit.  An output row IS
        // a thread here: thread r of the list publishes "row orow[r] is entry r" in shared memory (tagged with the epoch, two
        // maps alternate: one barrier per (net, tile), nothing to clear) and the owner adds the 20 values to its registers.
        // Load levels: (1) the records' fixed trailers, the first-tile row lists of all nets (both at fixed places of the record)
        // and the three base rows, (2) the values - the maps of all nets' first tiles are published behind ONE barrier.
        const int nets = p.comb_nets;
        const int rmax = p.comb_vcap / PPDE_Q;                                   // rows a net's list can hold (<= PAS_NT + 24)
        const int poff = p.comb_rec - PAS_MC32 - PAS_TRAIL - 2 * rmax;            // the (orow, cfirst) pairs sit at a fixed place
        uint4 tr[PAS_CN][2];
        int orow[PAS_CN];                                                        // my entry of every net's first-tile row list
#pragma unroll
        for (int k = 0; k < PAS_CN; ++k) {
            const uint16_t* rk = p.comb_wl + ((size_t)b * nets + (k < nets ? k : 0)) * p.comb_rec;
            const uint4* tp = reinterpret_cast<const uint4*>(rk + p.comb_rec - PAS_MC32 - PAS_TRAIL);
            tr[k][0] = (k < nets) ? __ldg(tp) : make_uint4(0u, 0u, 0u, 0u);
            tr[k][1] = (k < nets) ? __ldg(tp + 1) : make_uint4(0u, 0u, 0u, 0u);
            orow[k] = (k < nets && tid < rmax) ? (int)__ldg(rk + poff + 2 * tid) : 0;      // (garbage beyond the list: filtered below)
        }
        if (active) {
            load_row20(g, c.G + (int64_t)rx * NE + i * PPDE_Q);
            const int ip = i - m.win_lo;
            if (c.Gp && ip >= 0 && ip < m.Lp) {
                float a[PPDE_Q], d[PPDE_Q];
                load_row20(a, c.Gp + (int64_t)ry * m.D + ip * PPDE_Q);
                load_row20(d, c.Gp + (int64_t)rx * m.D + ip * PPDE_Q);
#pragma unroll
                for (int k = 0; k < PPDE_Q; ++k) g[k] += a[k] - d[k];
            }
        }
        // first tiles of all nets: one map per net, ONE barrier; entry = list index + 1 (0 = not an output row of this net)
#pragma unroll
        for (int k = 0; k < PAS_CN; ++k) s_map0[k][tid] = 0u;
        s_map[0][tid] = 0u; s_map[1][tid] = 0u;                // later tiles: tag 0 = no entry (epochs start at 1)
        __syncthreads();
        int r1k[PAS_CN], ntk[PAS_CN];
#pragma unroll
        for (int k = 0; k < PAS_CN; ++k) {
            ntk[k] = (int)(tr[k][0].x & 0xFFFFu);
            r1k[k] = ntk[k] ? (int)(tr[k][0].y & 0xFFFFu) : 0;                   // tstart[1]: rows of the first tile
            if (tid < r1k[k] && orow[k] < PAS_NT) s_map0[k][orow[k]] = (uint32_t)tid + 1u;
        }
        __syncthreads();
        uint32_t epoch = 0;
#pragma unroll
        for (int k = 0; k < PAS_CN; ++k) {
            if (k < nets) {
                const uint32_t tw[8] = {tr[k][0].x, tr[k][0].y, tr[k][0].z, tr[k][0].w, tr[k][1].x, tr[k][1].y, tr[k][1].z, tr[k][1].w};
                const uint16_t* pairs = p.comb_wl + ((size_t)b * nets + k) * p.comb_rec + poff;
                const float* v = p.comb_vals + ((size_t)k * n + b) * p.comb_vcap;
                {
                    const uint32_t me = active ? s_map0[k][i] : 0u;
                    if (me) {
                        float d[PPDE_Q];
                        load_row20(d, v + (size_t)(me - 1u) * PPDE_Q);
#pragma unroll
                        for (int q = 0; q < PPDE_Q; ++q) g[q] = fmaf(p.comb_scale, d[q], g[q]);
                    }
                    // (a first tile with more than PAS_NT rows cannot occur: <= 5 * 48 = 240 rows per tile)
                }
                int r0 = r1k[k];
                for (int tl = 1; tl < ntk[k]; ++tl) {                             // more than one tile of touched positions (12 % of the units)
                    const int wsel = (tl + 2) >> 1;                               // tstart[tl + 1] = 16-bit word tl + 2 of the trailer
                    uint32_t wv = tw[1];
#pragma unroll
                    for (int q = 2; q < 8; ++q) wv = (wsel == q) ? tw[q] : wv;
                    const int r1 = (tl + 2 < 2 * 8) ? (int)((wv >> (16 * (tl & 1))) & 0xFFFFu) : r0;
                    ++epoch;
                    uint32_t* map = s_map[epoch & 1u];
                    for (int r = r0 + tid; r < r1; r += NT) map[__ldg(pairs + 2 * r)] = (epoch << 16) | (uint32_t)(r - r0);
                    __syncthreads();
                    const uint32_t me = active ? map[i] : 0u;
                    if ((me >> 16) == epoch) {
                        float d[PPDE_Q];
                        load_row20(d, v + (size_t)(r0 + (int)(me & 0xFFFFu)) * PPDE_Q);
#pragma unroll
                        for (int q = 0; q < PPDE_Q; ++q) g[q] = fmaf(p.comb_scale, d[q], g[q]);
                    }
                    r0 = r1;
                }
            }
        }
        if (active) {
            float4* gy = reinterpret_cast<float4*>(c.G + (int64_t)ry * NE + i * PPDE_Q);
#pragma unroll
            for (int k = 0; k < 5; ++k) gy[k] = make_float4(g[4 * k], g[4 * k + 1], g[4 * k + 2], g[4 * k + 3]);
        }
    } else if (active) {
        load_row20(g, c.G + (int64_t)ry * NE + i * PPDE_Q);
    }
    float hgc = active ? -0.5f * pick20(g, zi) : -INFINITY;        // -0.5 G_y[i, z_i]; one position changes per sub-step
    float gmax = g[0];
#pragma unroll
    for (int k = 1; k < PPDE_Q; ++k) gmax = fmaxf(gmax, g[k]);

    for (int s = S_eff + tid; s < p.S; s += NT) c.lqr[(int64_t)s * n + b] = 0.f;
    for (int s = 0; s < S_eff; ++s) {
        const int64_t o = (int64_t)s * n + b;
        int cidx = cpre[0];
#pragma unroll
        for (int q = 1; q < 4; ++q) cidx = (s == q) ? cpre[q] : cidx;
        if (s >= 4) cidx = c.idx[o];
        const int pos = cidx / PPDE_Q, a = cidx - pos * PPDE_Q;
        if (s < U && i == pos) { zi = a; hgc = -0.5f * pick20(g, a); }        // state AFTER move s (ppde.py:124-125)
        float e[PPDE_Q];
        const float s3 = softmax_clamp_pos_fast<NW>(e, g, hgc, gmax, active, red);     // NO masks on the reverse path (ppde.py:126-127)
        if (i == pos) {
            const float lqr = logf(clamp_prob(pick20(e, a) / s3));
            c.lqr[o] = lqr;
            if (s < U) s_ratio += lqr - c.lqf[o];                  // u_mask * (rev - fwd), ppde.py:132 (one thread per sub-step, in order)
        }
    }
    __syncthreads();

    if (tid == 0) {
        const float log_ratio = s_ratio;
        const float e_x = pe_x, f_x = pf_x;
        const float e_y = pe_y, f_y = pf_y;
        const float log_acc = (e_y - e_x) + log_ratio;             // ppde.py:135-136
        const float u = u32_to_unit(Philox(p.seed)(0u, gid, (uint32_t)t, (uint32_t)(KIND_ACCEPT << 16)).x);
        const bool acc = expf(log_acc) >= u;                       // '>=' (ppde.py:138); NaN rejects
        c.log_acc[b] = log_acc;
        c.accept[b] = acc ? 1 : 0;
        const float e_rec = acc ? e_y : e_x, f_rec = acc ? f_y : f_x;   // ppde.py:141-143
        if (c.E_hist) c.E_hist[(int64_t)(t + 1) * n + b] = e_rec;
        if (c.fit_hist) c.fit_hist[(int64_t)(t + 1) * n + b] = f_rec;
        // 0: keep current state; 1: take y; 2: fall back to the paper-mode anchor
        s_flag = acc ? 1 : (p.paper_results ? 2 : 0);
        if (acc) { c.E[b] = e_y; c.fit[b] = f_y; c.row_cur[b] = ry; }
        else if (p.paper_results) {                                // x is never refreshed: reject = back to x0 (ppde.py:76-77,139)
            const int f = c.anchor_fixed ? c.anchor_fixed[b] : (c.row_wt - 2 * n);
            c.E[b] = c.E_fixed[f]; c.fit[b] = c.fit_fixed[f]; c.row_cur[b] = 2 * n + f;
        }
        const bool better = e_rec > pbest;                         // strict: first occurrence of the max (ppde.py:173)
        if (better) { c.best_E[b] = e_rec; c.best_fit[b] = f_rec; }
        s_flag |= better ? 4 : 0;
    }
    __syncthreads();
    const int flag = s_flag & 3;
    const bool better = (s_flag & 4) != 0;
    const uint8_t* src = ax;
    if (flag == 1) src = ay;
    else if (flag == 2) {
        const int f = c.anchor_fixed ? c.anchor_fixed[b] : (c.row_wt - 2 * n);
        src = c.aa_fixed + (int64_t)f * c.aa_stride;
    }
    // recorded state (post-accept, pre-reset): best-of-history and the random trajectory (ppde.py:142,146,172-183)
    int dpart = 0;
    uint8_t v = 0, wv = 0;
    if (active) {
        v = (uint8_t)(flag == 1 ? yi0 : (flag == 2 ? (int)src[i] : xi0)); wv = (uint8_t)wti;
        dpart = (v != wv);
        if (better) c.best_aa[(int64_t)b * c.aa_stride + i] = v;
        if (c.traj_aa && b == c.traj_chain) c.traj_aa[(int64_t)(t + 1) * c.aa_stride + i] = v;
    }
    const int dist = red_sum1i<NW>(dpart, redi);
    const bool reset = !p.paper_results && dist >= p.nmut_threshold;   // hard reset to WT (ppde.py:148-153)
    if (active) ax[i] = reset ? wv : v;
    if (reset && tid == 0) {
        const int f = c.row_wt - 2 * n;
        c.E[b] = c.E_fixed[f]; c.fit[b] = c.fit_fixed[f]; c.row_cur[b] = c.row_wt;
    }
}

// Known-answer entry point for the proposal arithmetic (tests only; not on the sampler's path): the SAME device functions the
// two kernels above use, driven by caller-provided logits.  Per row b:
//   dist[b]      = edit distance of aa[b] to wt                         (mut_distance, utils.py:5-14)
//   mask[b, e]   = 1 where mutation_mask masks entry e = i*20 + a       (utils.py:17-28, before `mask[~flag] = False`)
//   probs[b, :]  = clamp(softmax(logits[b] - LSE)) / sum               (safe_logits_to_probs + Categorical.__init__)
//   logp[b]      = log(clamp(probs[b, idx[b]]))                         (Categorical.log_prob)
template <int NT>
__global__ void __launch_bounds__(NT) pas_kat_kernel(const uint8_t* __restrict__ aa, int aa_stride, const uint8_t* __restrict__ wt, int L,
                                                     const float* __restrict__ logits, const int32_t* __restrict__ idx,
                                                     int32_t* __restrict__ dist, uint8_t* __restrict__ mask,
                                                     float* __restrict__ probs, float* __restrict__ logp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sP = reinterpret_cast<float*>(smem_raw);
    __shared__ float red[33];
    __shared__ int redi[33];
    const int b = blockIdx.x, NE = L * PPDE_Q, n4 = NE / 4;
    const uint8_t* a = aa + (int64_t)b * aa_stride;
    int dpart = 0;
    for (int i = threadIdx.x; i < L; i += NT) dpart += (a[i] != wt[i]);
    const int d = block_sum_int<NT>(dpart, redi);
    if (threadIdx.x == 0 && dist) dist[b] = d;
    if (mask)
        for (int e = threadIdx.x; e < NE; e += NT) {
            const int i = e / PPDE_Q, r = e - i * PPDE_Q;
            mask[(int64_t)b * NE + e] = (r != revert_only_target(a[i], wt[i])) ? 1 : 0;
        }
    if (!logits) return;
    float lmax = -INFINITY;
    for (int e = threadIdx.x; e < NE; e += NT) { const float l = logits[(int64_t)b * NE + e]; sP[e] = l; lmax = fmaxf(lmax, l); }
    __syncthreads();
    const float s3 = softmax_clamp_inplace<NT>(sP, n4, lmax, red);
    for (int e = threadIdx.x; e < NE; e += NT) probs[(int64_t)b * NE + e] = sP[e] / s3;
    if (threadIdx.x == 0 && logp && idx) logp[b] = logf(clamp_prob(sP[idx[b]] / s3));
}

}  // namespace ppde

using namespace ppde;

static size_t pas_smem(int L) { return (size_t)(2 * L * PPDE_Q + L) * sizeof(float) + 2 * ((L + 15) & ~15); }
// position-per-thread kernels (default for L <= PAS_NT); PPDE_PAS_POS=0 falls back to the strided kernels (A/B)
static bool pas_use_pos(int L) {
    static const bool on = [] { const char* e = getenv("PPDE_PAS_POS"); return !(e && e[0] == '0'); }();
    return on && L <= PAS_NT;
}
static int pas_check(const ppde_potts_t* m, const ppde_chains_t* c, const ppde_pas_params_t* p) {
    if (!m || !c || !p) return (int)cudaErrorInvalidValue;
    if (p->S < 1 || p->S > PPDE_MAX_S || c->L < 1 || c->aa_stride < c->L) return (int)cudaErrorInvalidValue;
    if (!c->aa || !c->aa_y || !c->row_cur || !c->G || !c->U || !c->idx || !c->old_aa || !c->lqf || !c->lqr || !m->wt)
        return (int)cudaErrorInvalidValue;
    return 0;
}

extern "C" int ppde_pas_propose(const ppde_potts_t* m, const ppde_chains_t* c, const ppde_pas_params_t* p,
                                void* stream) {
    if (int r = pas_check(m, c, p)) return r;
    if (c->n <= 0) return 0;
    if (p->fuse_potts && (!pas_use_pos(c->L) || !c->Gp || !c->Epotts_y || !m->Jsym || !m->h)) return (int)cudaErrorInvalidValue;
    if (pas_use_pos(c->L)) {
        if (p->uniforms) pas_propose_pos_kernel<4, true><<<c->n, PAS_NT, 0, (cudaStream_t)stream>>>(*m, *c, *p, PhiloxKeys(p->seed));
        else pas_propose_pos_kernel<4, false><<<c->n, PAS_NT, 0, (cudaStream_t)stream>>>(*m, *c, *p, PhiloxKeys(p->seed));
        return launch_done();
    }
    size_t smem = pas_smem(c->L);
    static SmemCache configured;
    if (cudaError_t e = ensure_dynamic_smem(pas_propose_kernel<128>, smem, configured)) return (int)e;
    pas_propose_kernel<128><<<c->n, 128, smem, (cudaStream_t)stream>>>(*m, *c, *p);
    return launch_done();
}

extern "C" int ppde_pas_reverse_accept(const ppde_potts_t* m, const ppde_chains_t* c, const ppde_pas_params_t* p,
                                       void* stream) {
    if (int r = pas_check(m, c, p)) return r;
    if (c->n <= 0) return 0;
    if (!c->E || !c->fit || !c->E_y || !c->fit_y || !c->log_acc || !c->accept || !c->best_E || !c->best_fit || !c->best_aa)
        return (int)cudaErrorInvalidValue;
    if (p->comb_nets > 0) {              // fused gradient combine: position-per-thread kernel only, every pointer it reads present
        if (!pas_use_pos(c->L) || p->comb_nets > PAS_CN || !p->comb_vals || !p->comb_wl || p->comb_vcap < 1 || p->comb_rec < 1)
            return (int)cudaErrorInvalidValue;
    }
    if (pas_use_pos(c->L)) {
        pas_reverse_accept_pos_kernel<<<c->n, PAS_NT, 0, (cudaStream_t)stream>>>(*m, *c, *p);
        return launch_done();
    }
    size_t smem = pas_smem(c->L);
    static SmemCache configured;
    if (cudaError_t e = ensure_dynamic_smem(pas_reverse_accept_kernel<128>, smem, configured)) return (int)e;
    pas_reverse_accept_kernel<128><<<c->n, 128, smem, (cudaStream_t)stream>>>(*m, *c, *p);
    return launch_done();
}

extern "C" int32_t ppde_pas_reverse_fuse_max_len(void) { return pas_use_pos(1) ? PAS_NT : 0; }

extern "C" int ppde_pas_kat(const uint8_t* aa, int32_t aa_stride, const uint8_t* wt, int32_t n, int32_t L, const float* logits,
                            const int32_t* idx, int32_t* dist, uint8_t* mask, float* probs, float* logp, void* stream) {
    if (n <= 0) return 0;
    if (!aa || !wt || L <= 0 || aa_stride < L || (logits && !probs)) return (int)cudaErrorInvalidValue;
    const size_t smem = (size_t)L * PPDE_Q * sizeof(float);
    static SmemCache configured;
    if (cudaError_t e = ensure_dynamic_smem(pas_kat_kernel<128>, smem, configured)) return (int)e;
    pas_kat_kernel<128><<<n, 128, smem, (cudaStream_t)stream>>>(aa, aa_stride, wt, L, logits, idx, dist, mask, probs, logp);
    return launch_done();
}
