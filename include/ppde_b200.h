/* ppde_b200 — C-ABI of the B200-native PPDE hot path (libppde_b200.so).
 *
 * The reference (pemami4911/ppde) has NO native boundary: its hot path is a duck-typed
 * Python protocol.  This header is the boundary a maintainer binds under
 * that protocol (ctypes stub in INTEGRATION.md).  Every entry point is a stateless launcher:
 * plain device pointers + sizes + a cudaStream_t (passed as void*), returns 0 or a
 * cudaError_t value (cudaErrorInvalidValue for NULL / inconsistent arguments), never aborts,
 * allocates nothing and keeps no state between calls except a launch counter and a per-device
 * cache of "dynamic shared memory limit already raised for this kernel" (profiling and tuning
 * switches are per-call arguments, ppde_tune_t).
 *
 * Reference interfaces each entry point replaces (paths relative to the reference root):
 *   ppde_potts_symmetrize        PottsModel.__init__ parameter load          ppde/nets.py:245-262
 *   ppde_potts_full              PottsModel.hamiltonian/forward + autograd   ppde/nets.py:282-299, ppde/energy.py:106-108
 *   ppde_potts_incremental       same quantity at y, reusing the field at x  
 *   ppde_potts_dense_full        same as ppde_potts_full as one tensor-core GEMM  ppde/nets.py:285-290 (the two einsums)
 *   ppde_cnn_forward             OnehotCNN.forward x3, EnsembleProtein mean  ppde/nets.py:363-376,434-442
 *   ppde_cnn_backward_combine    autograd through the CNN + PoE sum          ppde/energy.py:104-108
 *   ppde_cnn_dirty / ppde_cnn_forward_inc   the same forward for a proposal that differs from a cached state in a few
 *                                residues (sampler loop, energy at y)         ppde/protein_samplers/ppde.py:116-120, nets.py:363-376
 *   ppde_cnn_backward_tc[_rows] / ppde_cnn_backward_delta   the same gradient, exact or as (gradient at the current
 *                                state) + change                              ppde/energy.py:108
 *   ppde_pas_propose             PPDE_PAS.run forward path loop              ppde/protein_samplers/ppde.py:67-116
 *                                 + mut_distance / mutation_mask / safe_logits_to_probs  ppde/utils.py:5-28,106-111
 *   ppde_pas_reverse_accept      reverse proposal, MH accept, reset, history ppde/protein_samplers/ppde.py:122-153,172-183
 *   ppde_onehot_to_aa / ppde_aa_to_onehot   seqs_to_onehot / onehot2seq      ppde/third_party/hsu/data_utils.py:150-175
 *   ppde_population_metrics      mut_distance (n_hops) + sequence hash       ppde/utils.py:5-14, scripts/make_figures.py:29-36
 *   ppde_oracle_ridge            AugmentedLinearRegression.forward           ppde/nets.py:315-347 (log_every oracle call)
 *
 * Data layout (all row-major, device memory):
 *   residues     uint8  [n, aa_stride]   aa_stride >= L, multiple of 16; alphabet ACDEFGHIKLMNPQRSTVWY = 0..19
 *   gradient     float  [rows, 20*L]     entry (i,a) at i*20+a — same flattening as the reference's reshape(n,-1)
 *   Potts field  float  [rows, D]        D = 20*Lp, window positions only
 *   Jsym         float  [D, D]           Jsym[(i,a),(j,l)] = (J[i,j,a,l] + J[j,i,l,a]) / 2
 *   G / Gp pools: rows [0,n) and [n,2n) are the two private rows of chain b (b and n+b);
 *                 rows >= 2n are shared "fixed" rows (wild type, paper-mode anchors).
 */
#ifndef PPDE_B200_H
#define PPDE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PPDE_MAX_NETS 4
#define PPDE_MAX_SUBSTEPS 32

typedef struct ppde_potts {
    int32_t L;          /* full sequence length */
    int32_t Lp;         /* Potts window length (contiguous) */
    int32_t win_lo;     /* first window position, 0-based inclusive (index_list[0] - offset) */
    int32_t D;          /* 20 * Lp */
    const float* Jsym;  /* [D, D] */
    const float* h;     /* [D] */
    const uint8_t* wt;  /* [L] */
    float wt_H;         /* H(wt), subtracted from every energy (forward(delta=True)) */
    int32_t _pad;
} ppde_potts_t;

typedef struct ppde_cnn_net {
    const float* T0;    /* [5][20][C]   T0[t][a][c] = encoder.weight[c][a][t] */
    const float* b0;    /* [C]          encoder.bias */
    const float* W1;    /* [2C][C]      embedding.0.weight */
    const float* W1T;   /* [C][2C]      its transpose (forward K-major operand) */
    const float* b1;    /* [2C]         embedding.0.bias */
    const float* d;     /* [2C]         decoder.weight[0] */
    const float* W0r;   /* [C][5][20]   W0r[c][t][a] = encoder.weight[c][a][t] (backward) */
    const float* W1p;   /* [2C][kpad]   embedding.0.weight with rows zero-padded to kpad = roundup(C,16) floats */
    float c;            /* decoder.bias */
    float w1_scale;     /* power of two: max|W1| * w1_scale in [2^13, 2^14)  (fp16 operand split, tensor-core path) */
    float r1_scale;     /* power of two: (upper bound of r1) * r1_scale in [2^13, 2^14) */
    float w0_scale;     /* power of two for encoder.weight (tensor-core backward) */
    float adj_scale;    /* power of two for the adjoint rows sum_j d_j W1[j,c] (tensor-core backward) */
    int32_t _pad;
} ppde_cnn_net_t;

typedef struct ppde_cnn {
    int32_t n_nets;     /* 3 in the reference (EnsembleProtein) */
    int32_t C;          /* channels = L */
    int32_t L;
    int32_t P;          /* L - 4 conv output positions */
    ppde_cnn_net_t net[PPDE_MAX_NETS];
} ppde_cnn_t;

typedef struct ppde_chains {
    int32_t n;            /* local chains */
    int32_t chain_offset; /* global id of local chain 0 (random streams are indexed by global id) */
    int32_t L;
    int32_t aa_stride;
    uint8_t* aa;          /* [n, aa_stride] current state */
    uint8_t* aa_y;        /* [n, aa_stride] proposal y */
    int32_t* row_cur;     /* [n] pool row holding G / Gp of the current state */
    float* G;             /* pool [2n + n_fixed, 20L] */
    float* Gp;            /* pool [2n + n_fixed, D] */
    float* E;             /* [n] energy of current state */
    float* fit;           /* [n] CNN-ensemble fitness of current state */
    float* E_y;           /* [n] */
    float* fit_y;         /* [n] */
    float* Epotts_y;      /* [n] */
    /* fixed rows */
    int32_t n_fixed;
    int32_t row_wt;            /* pool row of the wild type (>= 2n) */
    const float* E_fixed;      /* [n_fixed] */
    const float* fit_fixed;    /* [n_fixed] */
    const uint8_t* aa_fixed;   /* [n_fixed, aa_stride] */
    const int32_t* anchor_fixed; /* [n] fixed-row index (0..n_fixed-1) of each chain's paper-mode anchor, or NULL */
    /* per-iteration trace (always written; tiny) */
    int32_t* U;           /* [n] path length */
    int32_t* idx;         /* [S, n] flat proposal index i*20+a */
    uint8_t* old_aa;      /* [S, n] residue at the proposed position before the move */
    float* lqf;           /* [S, n] forward log-prob */
    float* lqr;           /* [S, n] reverse log-prob */
    float* log_acc;       /* [n] */
    uint8_t* accept;      /* [n] */
    /* run-level outputs */
    float* E_hist;        /* [T+1, n] or NULL */
    float* fit_hist;      /* [T+1, n] or NULL */
    float* best_E;        /* [n] */
    float* best_fit;      /* [n] */
    uint8_t* best_aa;     /* [n, aa_stride] */
    uint8_t* traj_aa;     /* [T+1, aa_stride] states of chain `traj_chain`, or NULL */
    int32_t traj_chain;   /* local index, -1 if not on this rank */
    int32_t _pad;
} ppde_chains_t;

typedef struct ppde_pas_params {
    int32_t S;            /* sub-steps per iteration = 2*pas-1 (fixed; masked by U) */
    int32_t nmut_threshold; /* INT32_MAX when --nmut_threshold 0 */
    int32_t paper_results;
    int32_t t;            /* iteration index (stream counter) */
    int32_t min_pos;      /* proposals outside [min_pos, max_pos] are masked (run() arguments, ppde.py:59-63) */
    int32_t max_pos;      /* inclusive */
    uint64_t seed;
    const float* uniforms; /* optional materialised proposal uniforms [S, n, 20L]; NULL = Philox in-kernel */
    const int32_t* t_dev;  /* optional device-resident iteration counter (CUDA-graph replay); NULL = use t */
    int32_t full_trace;    /* 1: evaluate all S sub-steps of every chain (the reference computes, then masks, the sub-steps
                            * s >= U[b], ppde.py:83,111-115,132); 0: skip those dead sub-steps - nothing they produce reaches
                            * the state, the log-ratio or the accept decision - and record idx = -1, lqf = lqr = 0 for them */
    /* Fused gradient combine (ppde_pas_reverse_accept only; comb_nets = 0: off).  The reverse kernel stages the proposal's
     * gradient row in shared memory anyway (ppde.py:126-127 reads it for every reverse softmax); with comb_nets > 0 it first
     * ASSEMBLES that row - G[y] = G[x] + (Gp[y] - Gp[x])(window) + comb_scale * the sparse per-net changes
     * ppde_cnn_backward_delta left in its scratch when called with tune->parts = 3 (records + tensor-core kernel, no combine
     * kernel; layout from ppde_cnn_backward_delta_layout) - writes it to the pool once and goes on from shared memory:
     * ppde/energy.py:104-108 (one gradient, one sum) without a separate pass over four 19 KB rows per chain.
     * Needs L <= ppde_pas_reverse_fuse_max_len(), comb_nets <= 3, and c->Gp when the Potts expert is present. */
    int32_t comb_nets;
    const float* comb_vals;     /* [comb_nets][n][comb_vcap] values [row][20] of the output rows the records list */
    const uint16_t* comb_wl;    /* [n][comb_nets][comb_rec] records of cnn_delta_record_kernel */
    int32_t comb_vcap;
    int32_t comb_rec;
    float comb_scale;           /* lamda / n_nets */
    int32_t fuse_potts;         /* 1: ppde_pas_propose also does the work of ppde_potts_incremental for its chain (Gp of the
                                 * proposal = Gp of the current state + net coupling-row differences, Epotts_y from the new
                                 * field; nets.py:259-299): the CTA knows the moves, and the memory-bound row update of one
                                 * chain overlaps the Philox arithmetic of the others.  Same L limit as the fused combine. */
} ppde_pas_params_t;

/* Per-call tuning / measurement switches of the tensor-core CNN entry points (NULL = defaults = production behaviour).
 * No process-global state: two engines, streams or devices can use different settings concurrently. */
typedef struct ppde_tune {
    int32_t parts;        /* bit mask of the kernels a composite launcher runs (0 = all; per-kernel timing in bench.py):
                           * ppde_cnn_forward_inc: 1 scan, 2 tensor-core kernel, 4 merge;
                           * ppde_cnn_backward_tc[_rows] / ppde_cnn_backward_delta: 1 winner records, 2 tensor-core kernel, 4 combine */
    int32_t forward_ctas; /* ppde_cnn_forward_tc: 1 = one CTA per channel tile, 0 / 2 = CTA pairs with cta_group::2 MMAs */
    int32_t delta_layout; /* ppde_cnn_backward_delta: 0 = compact tiles of the touched positions, 1 = one column per position */
    int32_t dbg;          /* timing experiments only, results are WRONG with any bit set (tools/prof_*.py) */
    long long* prof;      /* non-NULL device buffer [grid][16] int64: instrumented build with per-role cycle counters */
} ppde_tune_t;

const char* ppde_version(void);
int ppde_last_launch_count(void);   /* kernels launched by this library since load (bench's gpu_launches) */

int ppde_potts_symmetrize(const float* J, int32_t Lp, float* Jsym, void* stream);
int ppde_potts_full(const ppde_potts_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n,
                    float* Gp, int64_t Gp_stride, float* Epotts, void* stream);
int ppde_potts_incremental(const ppde_potts_t* m, const ppde_chains_t* c, const ppde_pas_params_t* p, void* stream);
/* Dense re-evaluation on the tcgen05 tensor cores: Gp = h + onehot(aa) x Jsym as one [n x D] x [D x D] GEMM (same
 * quantity and contract as ppde_potts_full; any D).  `Jt` is a tiled fp16 hi/lo image of Jsym * jscale built once by
 * ppde_potts_dense_pack (ppde_potts_dense_image_bytes(D) bytes, 16-byte aligned); jscale = power of two with
 * max|Jsym| * jscale in [2^13, 2^14). */
int64_t ppde_potts_dense_image_bytes(int32_t D);
int ppde_potts_dense_pack(const ppde_potts_t* m, float jscale, void* Jt, void* stream);
int ppde_potts_dense_full(const ppde_potts_t* m, const void* Jt, float jscale, const uint8_t* aa, int32_t aa_stride,
                          int32_t n, float* Gp, int64_t Gp_stride, float* Epotts, void* stream);
int ppde_cnn_forward(const ppde_cnn_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n,
                     unsigned long long* mkey /* [n, n_nets, 2C] */, void* stream);
/* same contract as ppde_cnn_forward, on the tcgen05 tensor cores (needs C <= 256); writes every key, no memset.
 * r1mask (optional, [n, n_nets, P, 32] bytes): bit c of a position's 32 bytes = relu mask of the conv layer, consumed
 * by ppde_cnn_backward_tc. */
int ppde_cnn_forward_tc(const ppde_cnn_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n,
                        unsigned long long* mkey /* [n, n_nets, 2C] */, uint8_t* r1mask, const ppde_tune_t* tune,
                        void* stream);
int ppde_cnn_backward_combine(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa, int32_t aa_stride,
                              int32_t n, const unsigned long long* mkey, float lamda,
                              const float* Gp, int64_t Gp_stride, const int32_t* gp_rows,
                              const float* Epotts, float* G, int64_t G_stride, const int32_t* g_rows,
                              float* E, float* fit, void* stream);
/* gradient part of ppde_cnn_backward_combine on the tensor cores: G rows = Gp(window) + lamda/n_nets * sum_k dfit_k/dx.
 * Energies / fitness come from ppde_cnn_backward_combine(..., G = NULL, ...). Needs C <= 256. */
int ppde_cnn_backward_tc(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa, int32_t aa_stride,
                         int32_t n, const unsigned long long* mkey, float lamda,
                         const float* Gp, int64_t Gp_stride, const int32_t* gp_rows,
                         float* G, int64_t G_stride, const int32_t* g_rows,
                         const uint8_t* r1mask /* from ppde_cnn_forward_tc */,
                         float* scratch /* ppde_cnn_backward_scratch_floats(m, n) floats */,
                         const ppde_tune_t* tune, void* stream);
/* floats of `scratch` that ppde_cnn_backward_tc / _tc_rows / ppde_cnn_backward_delta need for n chains (per-net partial
 * gradients or sparse row values, followed by the winner records) */
int64_t ppde_cnn_backward_scratch_floats(const ppde_cnn_t* m, int32_t n);
/* same as ppde_cnn_backward_tc with the relu-mask rows taken from a POOL: chain b's mask lives in row
 * mask_rows[b] (NULL: mask_row_base + b) of r1mask [rows, n_nets, P, 32]. */
int ppde_cnn_backward_tc_rows(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa, int32_t aa_stride,
                              int32_t n, const unsigned long long* mkey, float lamda,
                              const float* Gp, int64_t Gp_stride, const int32_t* gp_rows,
                              float* G, int64_t G_stride, const int32_t* g_rows,
                              const uint8_t* r1mask, const int32_t* mask_rows, int32_t mask_row_base,
                              const int32_t* btab /* optional block table of the pools, see ppde_cnn_forward_inc */,
                              float* scratch, const ppde_tune_t* tune, void* stream);
/* Incremental CNN forward (same quantity as ppde_cnn_forward_tc, bit for bit; OnehotCNN.forward, ppde/nets.py:363-376).
 * A proposal differs from the chain's current state in a few residues, and a residue only moves the 5 conv rows that
 * read it, so the max-pool over positions is kept per BLOCK of PB positions in a pool
 *     bkey [rows, n_nets, NB = ceil(P/PB), 2C] uint64   (rows indexed like the G / Gp pools; PB = ppde_cnn_block_positions() = 8)
 * and only the dirty blocks of every chain are recomputed on the tensor cores.
 * A proposal row does not copy the clean blocks of the current state, it POINTS at them:
 *     btab [rows, NB] int32   btab[r][q] = the pool row whose slot holds block q of row r (its keys in bkey and its 16
 *                             relu-mask rows in r1mask [rows, n_nets, P, 32]); a fully evaluated row points at itself.
 *   ppde_cnn_dirty        dmask[b] = dirty-block bits of proposal aa_y[b] against the current state aa_x[b].
 *   ppde_cnn_forward_inc  recomputes the dirty blocks of aa[b] (dmask == NULL: all blocks = full evaluation) into free
 *                         slots, writes row rows_y[b] (NULL: row_base_y + b) of btab (clean blocks: the entries of row
 *                         rows_x[b]) and mkey [n, n_nets, 2C] as ppde_cnn_forward_tc does. */
int ppde_cnn_dirty(const ppde_cnn_t* m, const uint8_t* aa_x, const uint8_t* aa_y, int32_t aa_stride, int32_t n,
                   uint32_t* dmask /* [n] */, void* stream);
int64_t ppde_cnn_forward_inc_ws_bytes(int32_t n);   /* workspace of ppde_cnn_forward_inc (block prefix sums and lists) */
int32_t ppde_cnn_block_positions(void);              /* PB: conv-output positions per block of the max-pool cache (8) */
int ppde_cnn_forward_inc(const ppde_cnn_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n,
                         unsigned long long* mkey, uint8_t* r1mask, const uint32_t* dmask, unsigned long long* bkey,
                         int32_t* btab, const int32_t* rows_x, const int32_t* rows_y, int32_t row_base_y,
                         unsigned long long* mkey_pool /* optional [rows, n_nets, 2C][2]: the two largest raw block keys of every
                                                          pool row (winner, runner-up or 0); with it the merge reads this list and
                                                          the dirty blocks' keys instead of all NB keys per channel */,
                         void* ws /* ppde_cnn_forward_inc_ws_bytes(n) bytes */, const ppde_tune_t* tune, void* stream);
/* DELTA backward: the CNN part of the gradient changes between the current state x and the proposal y only through the
 * conv rows whose relu mask changed and the channels whose max-pool winner moved (a few percent of the winners), so
 *     G[rows_y[b]] = G[rows_x[b]] + (Gp[rows_y[b]] - Gp[rows_x[b]])(window) + lamda/n_nets * sum_k d(dfit_k/dx)
 * with only those adjoint rows gathered (same tensor-core kernel, signed winner records).  Rounding differences accumulate
 * (~1e-7 of max|G| per update): callers refresh with ppde_cnn_backward_tc_rows periodically.  mkey_pool [rows, n_nets, 2C][2]
 * holds the raw winners of every pool row (ppde_cnn_forward_inc), r1mask the relu-mask pool; same scratch as ppde_cnn_backward_tc. */
int ppde_cnn_backward_delta(const ppde_cnn_t* m, const ppde_potts_t* pm, const uint8_t* aa_x, const uint8_t* aa_y,
                            int32_t aa_stride, int32_t n, const unsigned long long* mkey_y,
                            const unsigned long long* mkey_pool, float lamda, const float* Gp, int64_t Gp_stride,
                            float* G, int64_t G_stride, const int32_t* rows_x, const int32_t* rows_y,
                            const uint8_t* r1mask, const int32_t* btab, float* scratch, const ppde_tune_t* tune,
                            void* stream);
/* Where ppde_cnn_backward_delta (compact records) keeps its outputs inside `scratch` for n chains: the sparse per-net
 * changes at scratch[0 .. n_nets * n * vcap) floats, the records (uint16, `rec` each) behind them at float offset wl_offset.
 * For ppde_pas_params_t.comb_* (fused combine in ppde_pas_reverse_accept). */
int ppde_cnn_backward_delta_layout(const ppde_cnn_t* m, int32_t n, int32_t* vcap, int32_t* rec, int64_t* wl_offset);
int32_t ppde_pas_reverse_fuse_max_len(void);   /* largest L for which ppde_pas_reverse_accept accepts comb_nets > 0 */
/* dH_potts of n states from field rows already in the pool: Epotts[b] = 1/2 sum_i (Gp[rows[b]][(i,aa_i)] + h) - H(wt);
 * rows == NULL means row b. */
int ppde_potts_energy_rows(const ppde_potts_t* m, const uint8_t* aa, int32_t aa_stride, int32_t n, const float* Gp,
                           int64_t Gp_stride, const int32_t* rows, float* Epotts, void* stream);
/* Oracle model = mean of the ridge heads over [sqrt(1/reg) dH_potts, sqrt(1/r_h) onehot]  (ppde/nets.py:315-347), heads
 * pre-averaged on the host: out[b] = sbar * dH[b] + sum_i wbar[20 i + aa_i] + cbar. */
int ppde_oracle_ridge(const float* wbar /* [20L] */, float sbar, float cbar, const uint8_t* aa, int32_t aa_stride, int32_t n,
                      int32_t L, const float* dH /* [n] or NULL */, float* out /* [n] */, void* stream);
int ppde_step_rows(const ppde_chains_t* c, int32_t* rows_y, void* stream);
int ppde_pas_propose(const ppde_potts_t* m, const ppde_chains_t* c, const ppde_pas_params_t* p, void* stream);
int ppde_pas_reverse_accept(const ppde_potts_t* m, const ppde_chains_t* c, const ppde_pas_params_t* p, void* stream);

int ppde_onehot_to_aa(const float* x, int32_t n, int32_t L, uint8_t* aa, int32_t aa_stride, void* stream);
int ppde_aa_to_onehot(const uint8_t* aa, int32_t aa_stride, int32_t n, int32_t L, float* x, void* stream);
/* the same two conversions on HOST buffers with `nthreads` host threads (no CUDA call): a host one-hot population is reduced
 * to 1 byte per residue before it crosses PCIe, and expanded after the copy back (80x fewer bytes than the float one-hot) */
int ppde_host_onehot_to_aa(const float* x, int64_t n, int32_t L, uint8_t* aa, int64_t aa_stride, int32_t nthreads);
int ppde_host_aa_to_onehot(const uint8_t* aa, int64_t aa_stride, int64_t n, int32_t L, float* x, int32_t nthreads);
int ppde_population_metrics(const uint8_t* aa, int32_t aa_stride, int32_t n, int32_t L, const uint8_t* wt,
                            int32_t* dist /* [n] */, unsigned long long* hash /* [n] */, void* stream);
int ppde_counter_add(int32_t* t_dev, int32_t inc, void* stream);

/* ---- log_every population report on the device (ppde/protein_samplers/ppde.py:155-170; scripts/make_figures.py:29-49;
 * torch.topk as in ppde/protein_samplers/cmaes.py:39).  Exact: radix select, integer sums, whole-sequence comparison.
 * Under torch.distributed the inputs are the all-gathered vectors; these kernels are single-GPU. */
/* out[j] = np.quantile(x[0..n), q[j]) (default 'linear' method, evaluated in double); x, q, out on the device */
int ppde_quantiles(const float* x, int64_t n, const double* q, int32_t nq, double* out, void* stream);
/* out = { sum accept, sum dist, sum dist^2, n } (int64, device); accept / dist may be NULL */
int ppde_population_sums(const uint8_t* accept, const int32_t* dist, int64_t n, long long* out, void* stream);
/* number of distinct sequences (diversity_score * K / 100): table = int32 workspace of ppde_unique_count_table_entries(n)
 * entries (a power of two >= 2n); count = one int32 on the device */
int64_t ppde_unique_count_table_entries(int64_t n);
int ppde_unique_count(const uint8_t* aa, int64_t aa_stride, int64_t n, int32_t L, int32_t* table, int64_t table_entries,
                      int32_t* count, void* stream);
/* torch.topk(x, k): k <= 1024 largest values, descending, ties by lowest id; id of element i = ids[i] (ids != NULL) or
 * index_base + i */
int ppde_topk(const float* x, int64_t n, int32_t k, int64_t index_base, const long long* ids, float* vals, long long* idx,
              void* stream);
/* out[j, :] = aa[idx[j] - index_base, :]   (sequences of the top-k chains) */
int ppde_gather_rows(const uint8_t* aa, int64_t aa_stride, const long long* idx, int32_t k, int64_t index_base, uint8_t* out,
                     void* stream);
/* Known-answer entry point of the proposal arithmetic (tests): the device functions of ppde_pas_propose /
 * ppde_pas_reverse_accept on caller-provided logits.  dist [n] (mut_distance, ppde/utils.py:5-14); mask u8 [n, 20L] (mutation_mask,
 * utils.py:17-28); probs [n, 20L] = clamp(softmax(logits - LSE)) / sum (utils.py:106-111 + Categorical.__init__);
 * logp [n] = log(clamp(probs))[idx] (Categorical.log_prob).  logits / idx / dist / mask / logp may be NULL. */
int ppde_pas_kat(const uint8_t* aa, int32_t aa_stride, const uint8_t* wt, int32_t n, int32_t L, const float* logits,
                 const int32_t* idx, int32_t* dist, uint8_t* mask, float* probs, float* logp, void* stream);

#ifdef __cplusplus
}
#endif
#endif
