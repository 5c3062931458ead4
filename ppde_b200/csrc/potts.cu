// Potts expert: symmetrised couplings, full field by row gather, incremental field update.
//
// Reference: PottsModel.hamiltonian / forward(delta=True)  ppde/nets.py:282-299
//            autograd of the same                            ppde/energy.py:106-108
// Closed forms (verified against the reference's autograd):
//   Gp[(j,l)] = dH/dx[(j,l)] = h[(j,l)] + sum_i Jsym[(i, aa_i), (j,l)]
//   H(x)      = 1/2 sum_i ( Gp[(i,aa_i)] + h[(i,aa_i)] )
//   Gp(y)     = Gp(x) + sum_m ( Jsym[(i_m,new_m), :] - Jsym[(i_m,old_m), :] )   (net changes m)
#include "common.cuh"
#include "../../include/ppde_b200.h"
#include "launch.cuh"

namespace ppde {

__global__ void potts_symmetrize_kernel(const float* __restrict__ J, int Lp, float* __restrict__ Jsym) {
    // J[i][j][k][l] -> M[(i,k)][(j,l)];  Jsym = (M + M^T)/2
    const int64_t D = (int64_t)Lp * PPDE_Q;
    const int64_t total = D * D;
    for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total;
         o += (int64_t)gridDim.x * blockDim.x) {
        int r = (int)(o / D), c = (int)(o % D);
        int i = r / PPDE_Q, k = r % PPDE_Q, j = c / PPDE_Q, l = c % PPDE_Q;
        float a = J[(((int64_t)i * Lp + j) * PPDE_Q + k) * PPDE_Q + l];
        float b = J[(((int64_t)j * Lp + i) * PPDE_Q + l) * PPDE_Q + k];
        Jsym[o] = 0.5f * (a + b);
    }
}

// One CTA per chain: Gp[b,:] = h + sum_i Jsym[row(i,aa_i), :], then the energy from the field.
template <int NT>
__global__ void __launch_bounds__(NT) potts_full_kernel(ppde_potts_t m, const uint8_t* __restrict__ aa,
                                                        int aa_stride, float* __restrict__ Gp,
                                                        int64_t Gp_stride, float* __restrict__ Epotts) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* srow = reinterpret_cast<int*>(smem_raw);          // [Lp] row index of each window position
    __shared__ float red[33];
    const int b = blockIdx.x;
    const uint8_t* a = aa + (int64_t)b * aa_stride + m.win_lo;
    for (int i = threadIdx.x; i < m.Lp; i += NT) srow[i] = i * PPDE_Q + a[i];
    __syncthreads();
    const int D4 = m.D / 4;
    const float4* J4 = reinterpret_cast<const float4*>(m.Jsym);
    const float4* h4 = reinterpret_cast<const float4*>(m.h);
    float4* out4 = reinterpret_cast<float4*>(Gp + (int64_t)b * Gp_stride);
    for (int c = threadIdx.x; c < D4; c += NT) {
        float4 acc = h4[c];
#pragma unroll 4
        for (int i = 0; i < m.Lp; ++i) {
            float4 v = __ldg(J4 + (int64_t)srow[i] * D4 + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        out4[c] = acc;
    }
    __syncthreads();
    if (Epotts) {
        const float* g = Gp + (int64_t)b * Gp_stride;
        float part = 0.f;
        for (int i = threadIdx.x; i < m.Lp; i += NT) part += g[srow[i]] + m.h[srow[i]];
        float tot = block_sum<NT>(part, red);
        if (threadIdx.x == 0) Epotts[b] = 0.5f * tot - m.wt_H;
    }
}

__device__ __forceinline__ int y_row_of(int row_cur, int b, int n) { return row_cur == b ? n + b : b; }

// One CTA per chain: Gp_y = Gp_x + net coupling-row differences; Epotts_y from the new field.
template <int NT>
__global__ void __launch_bounds__(NT) potts_incremental_kernel(ppde_potts_t m, ppde_chains_t c, int S) {
    __shared__ int s_new[PPDE_MAX_S], s_old[PPDE_MAX_S];
    __shared__ int s_cnt;
    __shared__ float red[33];
    const int b = blockIdx.x;
    const uint8_t* ax = c.aa + (int64_t)b * c.aa_stride;
    const uint8_t* ay = c.aa_y + (int64_t)b * c.aa_stride;
    if (threadIdx.x == 0) {
        int cnt = 0;
        const int U = c.U[b];
        for (int s = 0; s < S && s < U; ++s) {
            int pos = c.idx[(int64_t)s * c.n + b] / PPDE_Q;
            if (pos < m.win_lo || pos >= m.win_lo + m.Lp) continue;   // outside the Potts window: no coupling
            if (ax[pos] == ay[pos]) continue;                         // no net change at this position
            int rn = (pos - m.win_lo) * PPDE_Q + ay[pos];
            bool dup = false;
            for (int q = 0; q < cnt; ++q) dup |= (s_new[q] == rn);
            if (dup) continue;
            s_new[cnt] = rn;
            s_old[cnt] = (pos - m.win_lo) * PPDE_Q + ax[pos];
            ++cnt;
        }
        s_cnt = cnt;
    }
    __syncthreads();
    const int cnt = s_cnt;
    const int rx = c.row_cur[b];
    const int ry = y_row_of(rx, b, c.n);
    const int D4 = m.D / 4;
    const float4* J4 = reinterpret_cast<const float4*>(m.Jsym);
    const float4* gx = reinterpret_cast<const float4*>(c.Gp + (int64_t)rx * m.D);
    float4* gy = reinterpret_cast<float4*>(c.Gp + (int64_t)ry * m.D);
    for (int q = threadIdx.x; q < D4; q += NT) {
        float4 acc = gx[q];
        for (int k = 0; k < cnt; ++k) {
            float4 vn = __ldg(J4 + (int64_t)s_new[k] * D4 + q);
            float4 vo = __ldg(J4 + (int64_t)s_old[k] * D4 + q);
            acc.x += vn.x - vo.x; acc.y += vn.y - vo.y; acc.z += vn.z - vo.z; acc.w += vn.w - vo.w;
        }
        gy[q] = acc;
    }
    __syncthreads();
    const float* g = c.Gp + (int64_t)ry * m.D;
    float part = 0.f;
    for (int i = threadIdx.x; i < m.Lp; i += NT) {
        int r = i * PPDE_Q + ay[m.win_lo + i];
        part += g[r] + m.h[r];
    }
    float tot = block_sum<NT>(part, red);
    if (threadIdx.x == 0) c.Epotts_y[b] = 0.5f * tot - m.wt_H;
}

}  // namespace ppde

using namespace ppde;

extern "C" int ppde_potts_symmetrize(const float* J, int32_t Lp, float* Jsym, void* stream) {
    int64_t total = (int64_t)Lp * PPDE_Q * Lp * PPDE_Q;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    potts_symmetrize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(J, Lp, Jsym);
    return launch_done();
}

extern "C" int ppde_potts_full(const ppde_potts_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n,
                               float* Gp, int64_t Gp_stride, float* Epotts, void* stream) {
    if (n <= 0) return 0;
    if (m->D != m->Lp * PPDE_Q || (Gp_stride & 3)) return (int)cudaErrorInvalidValue;
    potts_full_kernel<256><<<n, 256, m->Lp * sizeof(int), (cudaStream_t)stream>>>(*m, aa, aa_stride, Gp, Gp_stride, Epotts);
    return launch_done();
}

extern "C" int ppde_potts_incremental(const ppde_potts_t* m, const ppde_chains_t* c, const ppde_pas_params_t* p,
                                      void* stream) {
    if (c->n <= 0) return 0;
    if (p->S > PPDE_MAX_S) return (int)cudaErrorInvalidValue;
    potts_incremental_kernel<256><<<c->n, 256, 0, (cudaStream_t)stream>>>(*m, *c, p->S);
    return launch_done();
}
