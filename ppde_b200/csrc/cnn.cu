// CNN-ensemble expert on integer residue states: forward (fp32 SIMT path) and closed-form backward,
// fused with the lambda-weighted product-of-experts sum.
//
// Reference: OnehotCNN.forward            ppde/nets.py:363-376
//            EnsembleProtein.__call__     ppde/nets.py:434-442 (mean of 3 nets, squeeze)
//            PoE energy + autograd        ppde/energy.py:104-108
// Closed forms (verified against autograd):
//   r1[p,c] = relu(b0[c] + sum_{t<5} W0[c, aa[p+t], t])                 one-hot conv = 5 gathers
//   r2[p,j] = relu(b1[j] + sum_c W1[j,c] r1[p,c]);  m[j] = max_p r2[p,j];  p*_j = lowest arg-max
//   fit_k   = c + sum_j d[j] m[j]
//   A[p,c]  = 1[r1[p,c]>0] * sum_{j: p*_j = p, m_j > 0} d[j] W1[j,c]
//   dfit_k/dx[i,a] = sum_{t<5} sum_c A[i-t,c] W0[c,a,t]
#include "common.cuh"
#include "../../include/ppde_b200.h"
#include "launch.cuh"

namespace ppde {

// ------------------------------------------------------------------ forward
// CTA = (position tile of TP, net k, chain b).  r1 tile kept in shared memory, W1^T streamed in
// K-chunks; 128x64 output tile per channel-tile iteration, 8x4 register tile per thread.
// Epilogue: bias + relu, max/arg-max over the tile's positions, one 64-bit atomicMax per channel:
// key = (float bits of r2 >= 0) << 32 | (0xFFFFFFFF - p)  -> larger value wins, ties -> lowest p.
constexpr int FW_TP = 64;        // positions per tile
constexpr int FW_TJ = 128;       // channels per tile
constexpr int FW_KC = 16;        // K chunk
constexpr int FW_NT = 256;
constexpr int FW_RS = FW_TP + 4; // r1 tile row stride (floats), keeps float4 alignment

__global__ void __launch_bounds__(FW_NT) cnn_forward_kernel(ppde_cnn_t m, const uint8_t* __restrict__ aa,
                                                            int aa_stride, unsigned long long* __restrict__ mkey) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = m.C, P = m.P, J2 = 2 * C;
    float* sR = reinterpret_cast<float*>(smem_raw);             // [C][FW_RS]  r1^T tile
    float* sW = sR + (size_t)C * FW_RS;                         // [FW_KC][FW_TJ] W1^T chunk
    __shared__ uint8_t sAA[FW_TP + 8];

    const int p0 = blockIdx.x * FW_TP, k = blockIdx.y, b = blockIdx.z;
    const ppde_cnn_net_t net = m.net[k];
    const uint8_t* a = aa + (int64_t)b * aa_stride;
    for (int i = threadIdx.x; i < FW_TP + 4; i += FW_NT) sAA[i] = (p0 + i < m.L) ? a[p0 + i] : 0;
    __syncthreads();
    // r1 tile: consecutive threads -> consecutive channels (coalesced T0 reads)
    for (int e = threadIdx.x; e < C * FW_TP; e += FW_NT) {
        const int pp = e / C, cc = e - pp * C;
        float v = 0.f;
        if (p0 + pp < P) {
            v = net.b0[cc];
#pragma unroll
            for (int t = 0; t < 5; ++t) v += __ldg(net.T0 + ((size_t)t * PPDE_Q + sAA[pp + t]) * C + cc);
            v = v > 0.f ? v : 0.f;
        }
        sR[(size_t)cc * FW_RS + pp] = v;
    }
    __syncthreads();

    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // tx: 4 positions, ty: 8 channels
    for (int j0 = 0; j0 < J2; j0 += FW_TJ) {
        float acc[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[r][q] = 0.f;
        for (int c0 = 0; c0 < C; c0 += FW_KC) {
            // W1T chunk [FW_KC][FW_TJ] <- W1T[c0+kk][j0+jj]  (W1T is [C][2C])
            for (int e = threadIdx.x; e < FW_KC * FW_TJ; e += FW_NT) {
                const int kk = e / FW_TJ, jj = e - kk * FW_TJ;
                float w = 0.f;
                if (c0 + kk < C && j0 + jj < J2) w = __ldg(net.W1T + (size_t)(c0 + kk) * J2 + j0 + jj);
                sW[e] = w;
            }
            __syncthreads();
            const int kmax = min(FW_KC, C - c0);
            for (int kk = 0; kk < kmax; ++kk) {
                const float4 w0 = *reinterpret_cast<const float4*>(sW + kk * FW_TJ + ty * 8);
                const float4 w1 = *reinterpret_cast<const float4*>(sW + kk * FW_TJ + ty * 8 + 4);
                const float4 r = *reinterpret_cast<const float4*>(sR + (size_t)(c0 + kk) * FW_RS + tx * 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                const float rv[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int rr = 0; rr < 8; ++rr)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[rr][q] = fmaf(wv[rr], rv[q], acc[rr][q]);
            }
            __syncthreads();
        }
        // epilogue
#pragma unroll
        for (int rr = 0; rr < 8; ++rr) {
            const int j = j0 + ty * 8 + rr;
            const float bias = (j < J2) ? net.b1[j] : 0.f;
            float best = -1.f; int bp = 0x7fffffff;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int p = p0 + tx * 4 + q;
                float v = acc[rr][q] + bias;
                v = v > 0.f ? v : 0.f;
                if (p < P && v > best) { best = v; bp = p; }
            }
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {                    // reduce over the 16 tx lanes
                const float ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int op = __shfl_xor_sync(0xffffffffu, bp, o);
                if (ov > best || (ov == best && op < bp)) { best = ov; bp = op; }
            }
            if (tx == 0 && j < J2 && best >= 0.f) {
                const unsigned long long key =
                    ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)bp);
                atomicMax(mkey + ((size_t)b * m.n_nets + k) * J2 + j, key);
            }
        }
    }
}

// ------------------------------------------------------------------ backward + PoE combine
// One CTA per chain, looping over the nets.  Shared memory:
//   sGc [20L]   accumulated sum_k dfit_k/dx      sKey [2C] winners      sCnt/sList: channels bucketed by p*
//   sA  [C][BW_PB] adjoint chunk                sY [BW_PB][100] conv-transpose chunk
constexpr int BW_NT = 256;
constexpr int BW_PB = 16;
constexpr int BW_AS = 20;       // adjoint chunk row stride (floats): 16B aligned, fewer bank conflicts

__global__ void __launch_bounds__(BW_NT) cnn_backward_combine_kernel(
    ppde_cnn_t m, ppde_potts_t pm, const uint8_t* __restrict__ aa, int aa_stride,
    const unsigned long long* __restrict__ mkey, float lamda,
    const float* __restrict__ Gp, int64_t Gp_stride, const int32_t* __restrict__ gp_rows,
    const float* __restrict__ Epotts, float* __restrict__ G, int64_t G_stride,
    const int32_t* __restrict__ g_rows, float* __restrict__ E, float* __restrict__ fit_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = m.C, P = m.P, L = m.L, J2 = 2 * C, NE = L * PPDE_Q;
    float* sGc = reinterpret_cast<float*>(smem_raw);            // [NE]
    float* sA = sGc + NE;                                       // [C][BW_PB]
    float* sY = sA + (size_t)C * BW_AS;                         // [BW_PB][100]
    float* sM = sY + BW_PB * 100;                               // [J2] d_j if winner counts else 0
    int* sPst = reinterpret_cast<int*>(sM + J2);                // [J2] p*_j
    int* sStart = sPst + J2;                                    // [P+1] bucket offsets
    int* sList = sStart + (P + 1);                              // [J2] channels sorted by p*
    int* sFill = sList + J2;                                    // [P]
    uint8_t* sAA = reinterpret_cast<uint8_t*>(sFill + P);       // [L]
    __shared__ float red[33];

    const int b = blockIdx.x;
    const uint8_t* a = aa + (int64_t)b * aa_stride;
    for (int i = threadIdx.x; i < L; i += BW_NT) sAA[i] = a[i];
    for (int i = threadIdx.x; i < NE; i += BW_NT) sGc[i] = 0.f;
    float fit_sum = 0.f;
    __syncthreads();

    for (int k = 0; k < m.n_nets; ++k) {
        const ppde_cnn_net_t net = m.net[k];
        const unsigned long long* keys = mkey + ((size_t)b * m.n_nets + k) * J2;
        for (int i = threadIdx.x; i <= P; i += BW_NT) sStart[i] = 0;
        for (int i = threadIdx.x; i < P; i += BW_NT) sFill[i] = 0;
        __syncthreads();
        float part = 0.f;
        for (int j = threadIdx.x; j < J2; j += BW_NT) {
            const unsigned long long key = keys[j];
            const float mj = __uint_as_float((unsigned)(key >> 32));
            const int pst = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFu));
            const float dj = net.d[j];
            part += dj * mj;
            const bool active = (mj > 0.f) && pst >= 0 && pst < P;   // relu'(0) = 0
            sM[j] = active ? dj : 0.f;
            sPst[j] = active ? pst : -1;
            if (active) atomicAdd(&sStart[pst + 1], 1);
        }
        const float dot = block_sum<BW_NT>(part, red);
        fit_sum += dot + net.c;
        if (threadIdx.x == 0) {                                   // exclusive scan (P <= ~1k: serial is fine)
            int run = 0;
            for (int i = 0; i <= P; ++i) { run += sStart[i]; sStart[i] = run; }
        }
        __syncthreads();
        // deterministic bucket fill: channel order ascending inside each bucket
        for (int pp = threadIdx.x; pp < P; pp += BW_NT) {
            if (sStart[pp + 1] == sStart[pp]) continue;
            int w = sStart[pp];
            for (int j = 0; j < J2; ++j) if (sPst[j] == pp) sList[w++] = j;
        }
        __syncthreads();

        for (int pb = 0; pb < P; pb += BW_PB) {
            const int npos = min(BW_PB, P - pb);
            if (sStart[min(pb + BW_PB, P)] == sStart[pb]) continue;   // no winners in this chunk (block-uniform)
            // adjoint chunk A[c][pp]
            for (int e = threadIdx.x; e < C * BW_PB; e += BW_NT) {
                const int pp = e / C, cc = e - pp * C;
                float acc = 0.f;
                if (pp < npos) {
                    const int p = pb + pp;
                    const int s0 = sStart[p], s1 = sStart[p + 1];
                    if (s1 > s0) {
                        float v = net.b0[cc];
#pragma unroll
                        for (int t = 0; t < 5; ++t) v += __ldg(net.T0 + ((size_t)t * PPDE_Q + sAA[p + t]) * C + cc);
                        if (v > 0.f)
                            for (int q = s0; q < s1; ++q) {
                                const int j = sList[q];
                                acc = fmaf(sM[j], __ldg(net.W1 + (size_t)j * C + cc), acc);
                            }
                    }
                }
                sA[(size_t)cc * BW_AS + pp] = acc;
            }
            __syncthreads();
            // Y[pp][ta] = sum_c A[c][pp] * W0r[c][ta]   (ta = t*20 + a), 4 positions per work item
            for (int wi = threadIdx.x; wi < 100 * (BW_PB / 4); wi += BW_NT) {
                const int ta = wi % 100, pg = wi / 100;
                float y0 = 0.f, y1 = 0.f, y2 = 0.f, y3 = 0.f;
                for (int cc = 0; cc < C; ++cc) {
                    const float w = __ldg(net.W0r + (size_t)cc * 100 + ta);
                    const float4 av = *reinterpret_cast<const float4*>(sA + (size_t)cc * BW_AS + pg * 4);
                    y0 = fmaf(av.x, w, y0); y1 = fmaf(av.y, w, y1); y2 = fmaf(av.z, w, y2); y3 = fmaf(av.w, w, y3);
                }
                sY[(pg * 4 + 0) * 100 + ta] = y0; sY[(pg * 4 + 1) * 100 + ta] = y1;
                sY[(pg * 4 + 2) * 100 + ta] = y2; sY[(pg * 4 + 3) * 100 + ta] = y3;
            }
            __syncthreads();
            // col2im, deterministic: each thread owns outputs (i,a), i in [pb, pb+npos+4)
            for (int e = threadIdx.x; e < (BW_PB + 4) * PPDE_Q; e += BW_NT) {
                const int di = e / PPDE_Q, aidx = e - di * PPDE_Q;
                const int i = pb + di;
                if (i >= L) continue;
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < 5; ++t) {
                    const int pp = di - t;
                    if (pp >= 0 && pp < npos) acc += sY[pp * 100 + t * PPDE_Q + aidx];
                }
                sGc[i * PPDE_Q + aidx] += acc;
            }
            __syncthreads();
        }
        __syncthreads();
    }

    const float fit = fit_sum / (float)m.n_nets;                 // torch.mean over the ensemble
    if (threadIdx.x == 0) {
        if (fit_out) fit_out[b] = fit;
        if (E) E[b] = (Epotts ? Epotts[b] : 0.f) + lamda * fit;  // energy.py:106-107
    }
    if (G) {
        const float scale = lamda / (float)m.n_nets;
        float* g = G + (int64_t)(g_rows ? g_rows[b] : b) * G_stride;
        const float* gp = Gp ? Gp + (int64_t)(gp_rows ? gp_rows[b] : b) * Gp_stride : nullptr;
        const int wlo = pm.win_lo * PPDE_Q, whi = (pm.win_lo + pm.Lp) * PPDE_Q;
        for (int j = threadIdx.x; j < NE; j += BW_NT) {
            const float pot = (gp && j >= wlo && j < whi) ? gp[j - wlo] : 0.f;
            g[j] = pot + scale * sGc[j];
        }
    }
}

// fitness only (get_energy path, no gradient): reads the winners, no backward.
__global__ void cnn_fit_kernel(ppde_cnn_t m, const unsigned long long* __restrict__ mkey, float lamda,
                               const float* __restrict__ Epotts, float* __restrict__ E,
                               float* __restrict__ fit_out, int n) {
    const int b = blockIdx.x;
    __shared__ float red[33];
    const int J2 = 2 * m.C;
    float fit_sum = 0.f;
    for (int k = 0; k < m.n_nets; ++k) {
        const unsigned long long* keys = mkey + ((size_t)b * m.n_nets + k) * J2;
        float part = 0.f;
        for (int j = threadIdx.x; j < J2; j += 128)
            part += m.net[k].d[j] * __uint_as_float((unsigned)(keys[j] >> 32));
        fit_sum += block_sum<128>(part, red) + m.net[k].c;
    }
    if (threadIdx.x == 0) {
        const float fit = fit_sum / (float)m.n_nets;
        if (fit_out) fit_out[b] = fit;
        if (E) E[b] = (Epotts ? Epotts[b] : 0.f) + lamda * fit;
    }
}

__global__ void step_rows_kernel(ppde_chains_t c, int32_t* rows_y) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < c.n) rows_y[b] = (c.row_cur[b] == b) ? c.n + b : b;
}

}  // namespace ppde

using namespace ppde;

extern "C" int ppde_cnn_forward(const ppde_cnn_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n,
                                unsigned long long* mkey, void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t keys = (size_t)n * m->n_nets * 2 * m->C;
    cudaError_t e = cudaMemsetAsync(mkey, 0, keys * sizeof(unsigned long long), st);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = ((size_t)m->C * FW_RS + FW_KC * FW_TJ) * sizeof(float);
    static SmemCache configured;
    e = ensure_dynamic_smem(cnn_forward_kernel, smem, configured);
    if (e != cudaSuccess) return (int)e;
    // grid.z is limited to 65535: split the chains over several launches if needed
    const int tiles = (m->P + FW_TP - 1) / FW_TP;
    for (int b0 = 0; b0 < n; b0 += 65535) {
        const int nb = (n - b0 < 65535) ? n - b0 : 65535;
        dim3 grid(tiles, m->n_nets, nb);
        cnn_forward_kernel<<<grid, FW_NT, smem, st>>>(*m, aa + (size_t)b0 * aa_stride, aa_stride,
                                                     mkey + (size_t)b0 * m->n_nets * 2 * m->C);
        int r = launch_done();
        if (r) return r;
    }
    return 0;
}

extern "C" int ppde_cnn_backward_combine(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa,
                                         int32_t aa_stride, int32_t n, const unsigned long long* mkey, float lamda,
                                         const float* Gp, int64_t Gp_stride, const int32_t* gp_rows,
                                         const float* Epotts, float* G, int64_t G_stride, const int32_t* g_rows,
                                         float* E, float* fit, void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (!G) {
        cnn_fit_kernel<<<n, 128, 0, st>>>(*m, mkey, lamda, Epotts, E, fit, n);
        return launch_done();
    }
    const int C = m->C, P = m->P, L = m->L, J2 = 2 * C;
    const size_t smem = ((size_t)L * PPDE_Q + (size_t)C * BW_AS + BW_PB * 100 + J2) * sizeof(float) +
                        ((size_t)J2 + (P + 1) + J2 + P) * sizeof(int) + ((L + 15) & ~15);
    static SmemCache configured;
    if (cudaError_t e = ensure_dynamic_smem(cnn_backward_combine_kernel, smem, configured)) return (int)e;
    cnn_backward_combine_kernel<<<n, BW_NT, smem, st>>>(*m, *pm, aa, aa_stride, mkey, lamda, Gp, Gp_stride, gp_rows,
                                                       Epotts, G, G_stride, g_rows, E, fit);
    return launch_done();
}

extern "C" int ppde_step_rows(const ppde_chains_t* c, int32_t* rows_y, void* stream) {
    if (c->n <= 0) return 0;
    step_rows_kernel<<<(c->n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*c, rows_y);
    return launch_done();
}
