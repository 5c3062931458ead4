"""Device-resident model + chain state and the per-iteration launch sequence.

Host mirror of the reference's hot loop (ppde/protein_samplers/ppde.py:65-153) and energy
(ppde/energy.py:97-108) on top of the C-ABI kernels.  PyTorch is used for device memory
and streams only; every arithmetic step is a kernel of libppde_b200.so.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import ChainsT, CnnT, PasParamsT, PottsT, TuneT

Q = 20
INT32_MAX = int(np.iinfo(np.int32).max)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("ppde_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device(device if device is not None else "cuda")


def aa_stride_for(L):
    return (L + 15) // 16 * 16


class Workspace:
    """Scratch buffers of the CNN kernels (winner keys, per-net gradient scratch + winner records, dirty-block lists, relu
    masks).  One per OWNER: the model has one for stand-alone evaluations (`PoEModel.energy`), every ChainEngine has its
    own, allocated once for the engine's n and kept alive with it - raw pointers into these buffers are baked into the
    engine's captured CUDA graphs, so they must never be reallocated or shared with a differently sized user."""

    def __init__(self, model):
        self.m = model
        self._buf = {}

    def _get(self, name, need, dtype):
        t = self._buf.get(name)
        if t is None or t.numel() < need:
            t = torch.empty(need, dtype=dtype, device=self.m.device)
            self._buf[name] = t
        return t

    def grad_scratch(self, n):
        m = self.m
        return self._get("gscratch", int(m.lib.ppde_cnn_backward_scratch_floats(C.byref(m.cnn), n)), torch.float32)

    def inc_ws(self, n):
        return self._get("inc_ws", int(self.m.lib.ppde_cnn_forward_inc_ws_bytes(n)), torch.uint8)

    def r1mask(self, n):
        m = self.m
        return self._get("r1mask", n * m.n_nets * m.P * 32, torch.uint8)

    def mkey(self, n):
        m = self.m
        return self._get("mkey", n * m.n_nets * 2 * m.C, torch.int64)


def _tune(parts=0, base=None):
    """ppde_tune_t for one call (None = production defaults)."""
    if not parts and base is None:
        return None
    t = TuneT(parts=int(parts))
    if base is not None:
        t.forward_ctas, t.delta_layout, t.dbg, t.prof = base.forward_ctas, base.delta_layout, base.dbg, base.prof
    return C.byref(t)


class PoEModel:
    """Product-of-experts weights resident in HBM (replicated on every GPU).

    J [Lp,Lp,20,20], h [Lp,20]  : potts.pkl 'J_ij', 'h_i' (ppde/nets.py:247-251)
    win_lo                      : index_list[0] - offset (ppde/nets.py:257-261)
    cnn                         : list of dicts W0[C,20,5] b0 W1[2C,C] b1 d[2C] c (OnehotCNN state, nets.py:350-361)
    lamda                       : --energy_lamda (ppde/energy.py:74)
    """

    def __init__(self, wt_aa, J, h, win_lo, cnn, lamda, device=None):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        dev = self.device
        wt_aa = np.asarray(wt_aa, dtype=np.uint8)
        self.L = int(wt_aa.shape[0])
        self.has_potts = J is not None          # False: ProteinSupervised (CNN-only, ppde/energy.py:143-164)
        self.Lp = int(J.shape[0]) if self.has_potts else 0
        self.win_lo = int(win_lo) if self.has_potts else 0
        self.D = Q * self.Lp
        self.NE = Q * self.L
        self.lamda = float(lamda)
        if self.win_lo < 0 or self.win_lo + self.Lp > self.L:
            raise ValueError("Potts window outside the sequence")
        self.aa_stride = aa_stride_for(self.L)
        with torch.cuda.device(dev):
            self.Jsym = torch.empty(self.D, self.D, dtype=torch.float32, device=dev)
            self.h = torch.zeros(max(self.D, 1), dtype=torch.float32, device=dev)
            if self.has_potts:
                if tuple(J.shape) != (self.Lp, self.Lp, Q, Q) or tuple(np.shape(h)) != (self.Lp, Q):
                    raise ValueError(f"J must be [Lp,Lp,20,20] and h [Lp,20]; got {tuple(J.shape)}, {tuple(np.shape(h))}")
                Jd = torch.as_tensor(np.ascontiguousarray(J, dtype=np.float32)).to(dev)
                _lib.check(self.lib.ppde_potts_symmetrize(_ptr(Jd), self.Lp, _ptr(self.Jsym), _stream()), "potts_symmetrize")
                torch.cuda.current_stream().synchronize()
                del Jd
                self.h = torch.as_tensor(np.ascontiguousarray(h, dtype=np.float32).reshape(-1)).to(dev)
            wt_pad = np.zeros(self.aa_stride, dtype=np.uint8)
            wt_pad[:self.L] = wt_aa
            self.wt = torch.from_numpy(wt_pad).to(dev)
            self.wt_host = wt_aa.copy()
            self.potts = PottsT(L=self.L, Lp=self.Lp, win_lo=self.win_lo, D=self.D, Jsym=self.Jsym.data_ptr(),
                                h=self.h.data_ptr(), wt=self.wt.data_ptr(), wt_H=0.0)
            # H(wt): same kernel, same arithmetic as every later evaluation (ppde/nets.py:262)
            self.wt_H = 0.0
            if self.has_potts:
                gp = torch.empty(1, self.D, dtype=torch.float32, device=dev)
                ep = torch.empty(1, dtype=torch.float32, device=dev)
                _lib.check(self.lib.ppde_potts_full(C.byref(self.potts), _ptr(self.wt), self.aa_stride, 1, _ptr(gp),
                                                    self.D, _ptr(ep), _stream()), "potts_full(wt)")
                self.wt_H = float(ep.item())
                self.potts.wt_H = self.wt_H
                # dense tensor-core path for bulk re-evaluation: tiled fp16 hi/lo image of Jsym * 2^k (built once)
                jmax = float(self.Jsym.abs().max().item())
                self.jscale = 2.0 ** (13 - int(np.ceil(np.log2(max(jmax, 1e-30)))))
                nbytes = int(self.lib.ppde_potts_dense_image_bytes(self.D))
                self.Jt = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                _lib.check(self.lib.ppde_potts_dense_pack(C.byref(self.potts), self.jscale, _ptr(self.Jt), _stream()),
                           "potts_dense_pack")
            # CNN ensemble
            self.n_nets = len(cnn)
            if self.n_nets > _lib.MAX_NETS:
                raise ValueError("too many ensemble members")
            self.C = self.L
            self.P = self.L - 4
            self._cnn_keep = []
            self.cnn = CnnT(n_nets=self.n_nets, C=self.C, L=self.L, P=self.P)
            for k, net in enumerate(cnn):
                W0 = torch.as_tensor(np.asarray(net["W0"], dtype=np.float32))
                if tuple(W0.shape) != (self.C, Q, 5):
                    raise ValueError(f"encoder.weight must be [{self.C},20,5], got {tuple(W0.shape)}")
                W1 = torch.as_tensor(np.asarray(net["W1"], dtype=np.float32))
                t = {
                    "T0": W0.permute(2, 1, 0).contiguous(),            # [5][20][C]
                    "b0": torch.as_tensor(np.asarray(net["b0"], dtype=np.float32)),
                    "W1": W1.contiguous(),
                    "W1T": W1.t().contiguous(),
                    "b1": torch.as_tensor(np.asarray(net["b1"], dtype=np.float32)),
                    "d": torch.as_tensor(np.asarray(net["d"], dtype=np.float32).reshape(-1)),
                    "W0r": W0.permute(0, 2, 1).contiguous(),           # [C][5][20]
                }
                kpad = (self.C + 15) // 16 * 16
                W1p = torch.zeros(2 * self.C, kpad, dtype=torch.float32)
                W1p[:, :self.C] = W1
                t["W1p"] = W1p
                t = {kk: v.to(dev) for kk, v in t.items()}
                self._cnn_keep.append(t)
                cn = self.cnn.net[k]
                for kk, v in t.items():
                    setattr(cn, kk, v.data_ptr())
                cn.c = float(np.asarray(net["c"]).reshape(-1)[0])
                # power-of-two operand scales for the fp16 hi/lo split of the tensor-core path
                w1max = float(W1.abs().max())
                r1max = float((torch.as_tensor(np.asarray(net["b0"], dtype=np.float32))
                               + W0.amax(dim=1).clamp_min(0).sum(dim=1)).clamp_min(0).max())   # bound on relu(conv)
                cn.w1_scale = 2.0 ** (13 - int(np.ceil(np.log2(max(w1max, 1e-30)))))
                cn.r1_scale = 2.0 ** (13 - int(np.ceil(np.log2(max(r1max, 1e-30)))))
                dvec = torch.as_tensor(np.asarray(net["d"], dtype=np.float32).reshape(-1))
                adjmax = float((dvec.abs()[:, None] * W1.abs()).sum(0).max())      # bound on |sum_j d_j W1[j,c]|
                cn.w0_scale = 2.0 ** (13 - int(np.ceil(np.log2(max(float(W0.abs().max()), 1e-30)))))
                cn.adj_scale = 2.0 ** (13 - int(np.ceil(np.log2(max(adjmax, 1e-30)))))
        self.ws = Workspace(self)             # scratch of stand-alone evaluations (engines own theirs)
        # A/B switches of the tensor-core kernels, per call (ppde_tune_t): PPDE_TC_CTAS=1 = 1-CTA forward kernel,
        # PPDE_BWD_DELTA_COMPACT=0 = one column per position in the delta backward
        self.tune = None
        dbg = int(os.environ.get("PPDE_TUNE_DBG", "0"))       # 8: the merge scans all NB keys (A/B of the top-2 list)
        if os.environ.get("PPDE_TC_CTAS", "2") == "1" or os.environ.get("PPDE_BWD_DELTA_COMPACT", "1") == "0" or dbg:
            self.tune = TuneT(forward_ctas=1 if os.environ.get("PPDE_TC_CTAS", "2") == "1" else 0,
                              delta_layout=1 if os.environ.get("PPDE_BWD_DELTA_COMPACT", "1") == "0" else 0, dbg=dbg)
        # CNN forward implementation: tcgen05 tensor-core kernel (needs the W1 tile in 256 TMEM columns,
        # i.e. C <= 256) or the fp32 SIMT kernel.  PPDE_CNN_FORWARD=simt forces the latter (A/B tests).
        want = os.environ.get("PPDE_CNN_FORWARD", "tc")
        self.cnn_forward_impl = "tc" if (want == "tc" and self.C <= 256) else "simt"
        wantb = os.environ.get("PPDE_CNN_BACKWARD", "tc")
        self.cnn_backward_impl = "tc" if (wantb == "tc" and self.C <= 256) else "simt"
        # incremental CNN forward (block-wise max-pool cache in the chain engine's row pools); needs both tensor-core
        # kernels and <= 16 blocks of 16 positions.  PPDE_CNN_INC=0 re-evaluates every position of every proposal.
        self.PB = int(self.lib.ppde_cnn_block_positions())              # positions per block of the max-pool cache
        self.NB = (self.P + self.PB - 1) // self.PB
        self.cnn_inc = (os.environ.get("PPDE_CNN_INC", "1") != "0" and self.cnn_forward_impl == "tc"
                        and self.cnn_backward_impl == "tc" and self.NB <= 32)
        # delta backward on top of the incremental forward: gradient rows updated by the change of the few adjoint rows that
        # differ between the current state and the proposal; an exact (full) backward every `bwd_refresh` iterations bounds the
        # accumulated rounding.  PPDE_CNN_BWD_DELTA=0 always runs the full backward.
        self.cnn_bwd_delta = self.cnn_inc and os.environ.get("PPDE_CNN_BWD_DELTA", "1") != "0"
        self.bwd_refresh = max(1, int(os.environ.get("PPDE_BWD_REFRESH", "32")))
        # full Potts evaluation: "dense" = tcgen05 GEMM for batches of >= dense_min chains, "gather" = row-gather kernel
        self.potts_full_impl = os.environ.get("PPDE_POTTS_FULL", "dense")
        self.dense_min = int(os.environ.get("PPDE_POTTS_DENSE_MIN", "512"))

    def cnn_forward(self, aa, n, mk, st, ws=None):
        ws = ws or self.ws
        if self.cnn_forward_impl == "tc":
            rm = _ptr(ws.r1mask(n)) if self.cnn_backward_impl == "tc" else C.c_void_p(0)
            _lib.check(self.lib.ppde_cnn_forward_tc(C.byref(self.cnn), _ptr(aa), self.aa_stride, n, _ptr(mk), rm,
                                                    _tune(0, self.tune), st), "cnn_forward_tc")
        else:
            _lib.check(self.lib.ppde_cnn_forward(C.byref(self.cnn), _ptr(aa), self.aa_stride, n, _ptr(mk), st),
                       "cnn_forward")

    def cnn_backward_combine(self, aa, n, mk, gp_ptr, gp_rows, ep_ptr, g_ptr, g_rows, E, fit, st, ws=None):
        """fit / E from the winners, then (if g_ptr) the gradient rows G = Gp(window) + lamda/n_nets * sum_k dfit_k/dx."""
        ws = ws or self.ws
        lib = self.lib
        null = C.c_void_p(0)
        if self.cnn_backward_impl == "tc" and self.cnn_forward_impl == "tc" and g_ptr:
            _lib.check(lib.ppde_cnn_backward_combine(
                C.byref(self.cnn), C.byref(self.potts), _ptr(aa), self.aa_stride, n, _ptr(mk), self.lamda,
                null, self.D, null, ep_ptr, null, self.NE, null, _ptr(E), _ptr(fit), st), "cnn_fit")
            _lib.check(lib.ppde_cnn_backward_tc(
                C.byref(self.cnn), C.byref(self.potts), _ptr(aa), self.aa_stride, n, _ptr(mk), self.lamda,
                gp_ptr, self.D, gp_rows, g_ptr, self.NE, g_rows, _ptr(ws.r1mask(n)), _ptr(ws.grad_scratch(n)),
                _tune(0, self.tune), st), "cnn_backward_tc")
        else:
            _lib.check(lib.ppde_cnn_backward_combine(
                C.byref(self.cnn), C.byref(self.potts), _ptr(aa), self.aa_stride, n, _ptr(mk), self.lamda,
                gp_ptr, self.D, gp_rows, ep_ptr, g_ptr, self.NE, g_rows, _ptr(E), _ptr(fit), st), "cnn_backward_combine")

    # -- CNN with the block-key / relu-mask POOLS of a chain engine (incremental path) -----------------
    def cnn_forward_pool(self, aa, n, mk, bkey, r1pool, dmask, rows_x, rows_y, row_base_y, st, mkpool=None, btab=None,
                         ws=None, parts=0):
        """Forward of n states into pool rows rows_y (None: row_base_y + b).  dmask None: every block is evaluated;
        otherwise only the dirty blocks, the others come from rows_x (ppde_cnn_forward_inc).  mkpool: the pool of RAW row
        winners (lets the merge read 1 + #dirty keys per channel instead of all NB)."""
        ws = ws or self.ws
        _lib.check(self.lib.ppde_cnn_forward_inc(C.byref(self.cnn), _ptr(aa), self.aa_stride, n, _ptr(mk), _ptr(r1pool),
                                                 _ptr(dmask), _ptr(bkey), _ptr(btab), _ptr(rows_x), _ptr(rows_y), int(row_base_y),
                                                 _ptr(mkpool), _ptr(ws.inc_ws(n)), _tune(parts, self.tune), st), "cnn_forward_inc")

    def cnn_backward_pool(self, aa, n, mk, gp_ptr, gp_rows, ep_ptr, g_ptr, g_rows, E, fit, r1pool, mask_rows, mask_row_base, st,
                          do_fit=True, do_grad=True, btab=None, ws=None, parts=0):
        ws = ws or self.ws
        lib = self.lib
        null = C.c_void_p(0)
        if do_fit:
            _lib.check(lib.ppde_cnn_backward_combine(
                C.byref(self.cnn), C.byref(self.potts), _ptr(aa), self.aa_stride, n, _ptr(mk), self.lamda,
                null, self.D, null, ep_ptr, null, self.NE, null, _ptr(E), _ptr(fit), st), "cnn_fit")
        if do_grad:
            _lib.check(lib.ppde_cnn_backward_tc_rows(
                C.byref(self.cnn), C.byref(self.potts), _ptr(aa), self.aa_stride, n, _ptr(mk), self.lamda,
                gp_ptr, self.D, gp_rows, g_ptr, self.NE, g_rows, _ptr(r1pool), _ptr(mask_rows), int(mask_row_base), _ptr(btab),
                _ptr(ws.grad_scratch(n)), _tune(parts, self.tune), st), "cnn_backward_tc_rows")

    def cnn_backward_delta(self, aa_x, aa_y, n, mk, mkpool, gp_ptr, ep_ptr, g_ptr, rows_x, rows_y, E, fit, r1pool, st,
                           do_fit=True, do_grad=True, btab=None, ws=None, parts=0):
        """fit / E of the proposals, and their gradient rows as  G[rows_y] = G[rows_x] + change  (ppde_cnn_backward_delta)."""
        ws = ws or self.ws
        lib = self.lib
        null = C.c_void_p(0)
        if do_fit:
            _lib.check(lib.ppde_cnn_backward_combine(
                C.byref(self.cnn), C.byref(self.potts), _ptr(aa_y), self.aa_stride, n, _ptr(mk), self.lamda,
                null, self.D, null, ep_ptr, null, self.NE, null, _ptr(E), _ptr(fit), st), "cnn_fit")
        if do_grad:
            _lib.check(lib.ppde_cnn_backward_delta(
                C.byref(self.cnn), C.byref(self.potts), _ptr(aa_x), _ptr(aa_y), self.aa_stride, n, _ptr(mk), _ptr(mkpool),
                self.lamda, gp_ptr, self.D, g_ptr, self.NE, _ptr(rows_x), _ptr(rows_y), _ptr(r1pool), _ptr(btab),
                _ptr(ws.grad_scratch(n)), _tune(parts, self.tune), st), "cnn_backward_delta")

    # -- full evaluation ------------------------------------------------------------------------
    def evaluate_into(self, aa, n, G, g_row0, Gp, gp_row0, E, fit, Epotts, want_grad=True, bkey=None, r1pool=None, mkpool=None,
                      btab=None, ws=None):
        """Energy (+ gradient field) of n states `aa` [n, aa_stride] written into pool rows
        g_row0.. / gp_row0.. (contiguous). E, fit, Epotts: float tensors [n].
        bkey / r1pool: the engine's block-key and relu-mask pools (rows g_row0.. are filled too)."""
        lib = self.lib
        st = _stream()
        gp_ptr, ep_ptr = C.c_void_p(0), C.c_void_p(0)
        if self.has_potts:
            gp_ptr, ep_ptr = C.c_void_p(Gp.data_ptr() + gp_row0 * self.D * 4), _ptr(Epotts)
            self.potts_full(aa, n, gp_ptr, ep_ptr, st)
        ws = ws or self.ws
        mk = ws.mkey(n)
        g_ptr = C.c_void_p(G.data_ptr() + g_row0 * self.NE * 4) if want_grad else C.c_void_p(0)
        if bkey is not None and want_grad:
            self.cnn_forward_pool(aa, n, mk, bkey, r1pool, None, None, None, g_row0, st, mkpool=mkpool, btab=btab, ws=ws)
            self.cnn_backward_pool(aa, n, mk, gp_ptr, C.c_void_p(0), ep_ptr, g_ptr, C.c_void_p(0), E, fit,
                                   r1pool, None, g_row0, st, btab=btab, ws=ws)
            return
        self.cnn_forward(aa, n, mk, st, ws=ws)
        self.cnn_backward_combine(aa, n, mk, gp_ptr, C.c_void_p(0), ep_ptr, g_ptr if want_grad else None,
                                  C.c_void_p(0), E, fit, st, ws=ws)

    def potts_full(self, aa, n, gp_ptr, ep_ptr, st, impl=None):
        """Potts field rows (contiguous, stride D) + energies of n states: dense tensor-core GEMM for large batches,
        row gather otherwise (ppde/nets.py:282-299 + autograd)."""
        impl = impl or (self.potts_full_impl if n >= self.dense_min else "gather")
        if impl == "dense":
            _lib.check(self.lib.ppde_potts_dense_full(C.byref(self.potts), _ptr(self.Jt), self.jscale, _ptr(aa),
                                                      self.aa_stride, n, gp_ptr, self.D, ep_ptr, st), "potts_dense_full")
        else:
            _lib.check(self.lib.ppde_potts_full(C.byref(self.potts), _ptr(aa), self.aa_stride, n, gp_ptr, self.D,
                                                ep_ptr, st), "potts_full")

    def onehot_to_aa(self, x):
        n, L, q = x.shape
        if L != self.L or q != Q:
            raise ValueError(f"expected one-hot [n,{self.L},20], got {tuple(x.shape)}")
        x = x.detach().to(self.device, torch.float32).contiguous()
        aa = torch.zeros(n, self.aa_stride, dtype=torch.uint8, device=self.device)
        _lib.check(self.lib.ppde_onehot_to_aa(_ptr(x), n, L, _ptr(aa), self.aa_stride, _stream()), "onehot_to_aa")
        return aa

    def aa_to_onehot(self, aa, n=None):
        n = aa.shape[0] if n is None else n
        x = torch.empty(n, self.L, Q, dtype=torch.float32, device=self.device)
        _lib.check(self.lib.ppde_aa_to_onehot(_ptr(aa), self.aa_stride, n, self.L, _ptr(x), _stream()), "aa_to_onehot")
        return x

    def energy(self, aa, want_grad=True):
        """(E [n], fit [n], G [n,L,20] or None, Epotts [n]) for residue states aa [n, aa_stride] on device."""
        n = aa.shape[0]
        dev = self.device
        E = torch.empty(n, dtype=torch.float32, device=dev)
        fit = torch.empty(n, dtype=torch.float32, device=dev)
        Ep = torch.zeros(n, dtype=torch.float32, device=dev)
        Gp = torch.empty(n, self.D, dtype=torch.float32, device=dev)
        G = torch.empty(n, self.NE, dtype=torch.float32, device=dev) if want_grad else None
        self.evaluate_into(aa, n, G, 0, Gp, 0, E, fit, Ep, want_grad)
        return E, fit, (G.view(n, self.L, Q) if want_grad else None), Ep


class ChainEngine:
    """n local chains of the PPDE sampler (one engine per GPU / rank)."""

    def __init__(self, model: PoEModel, n, pas_length=2, nmut_threshold=0, paper_results=False, seed=0,
                 chain_offset=0, num_steps=None, traj_chain=-1, min_pos=None, max_pos=None):
        self.m = model
        self.lib = model.lib
        self.n = int(n)
        self.S = 2 * int(pas_length) - 1
        if not (1 <= self.S <= _lib.MAX_SUBSTEPS):
            raise ValueError(f"ppde_pas_length must give 1 <= 2*pas-1 <= {_lib.MAX_SUBSTEPS}")
        self.thr = int(nmut_threshold) if nmut_threshold else INT32_MAX      # ppde.py:15-17
        self.paper = bool(paper_results)
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.chain_offset = int(chain_offset)
        self.T = num_steps
        self.t = 0
        self.traj_chain = int(traj_chain)
        # proposal window = run()'s (min_pos, max_pos), inclusive (ppde.py:59-63); default: the Potts window
        self.min_pos = int(min_pos) if min_pos is not None else (model.win_lo if model.has_potts else 0)
        self.max_pos = int(max_pos) if max_pos is not None else (
            model.win_lo + model.Lp - 1 if model.has_potts else model.L - 1)
        self._allocated = False
        self._graph = None
        self.ws = Workspace(model)             # this engine's own kernel scratch (its pointers live in the captured graphs)
        # True: also evaluate the sub-steps s >= U[b] that the reference computes and then masks (complete idx / lqf / lqr
        # trace, as the golden comparisons read it); False: skip them (same states, energies and accept decisions)
        self.full_trace = False

    # -- allocation ------------------------------------------------------------------------------
    def _alloc(self, n_fixed):
        m, n, dev = self.m, self.n, self.m.device
        f32, i32, u8 = torch.float32, torch.int32, torch.uint8
        S = self.S
        self.n_fixed = n_fixed
        rows = 2 * n + n_fixed
        self.G = torch.empty(rows, m.NE, dtype=f32, device=dev)
        self.Gp = torch.empty(rows, m.D, dtype=f32, device=dev)
        self.aa = torch.zeros(n, m.aa_stride, dtype=u8, device=dev)
        self.aa_y = torch.zeros(n, m.aa_stride, dtype=u8, device=dev)
        self.row_cur = torch.zeros(n, dtype=i32, device=dev)
        self.rows_y = torch.zeros(n, dtype=i32, device=dev)
        self.E = torch.zeros(n, dtype=f32, device=dev)
        self.fit = torch.zeros(n, dtype=f32, device=dev)
        self.E_y = torch.zeros(n, dtype=f32, device=dev)
        self.fit_y = torch.zeros(n, dtype=f32, device=dev)
        self.Epotts_y = torch.zeros(n, dtype=f32, device=dev)
        self.E_fixed = torch.zeros(n_fixed, dtype=f32, device=dev)
        self.fit_fixed = torch.zeros(n_fixed, dtype=f32, device=dev)
        self.aa_fixed = torch.zeros(n_fixed, m.aa_stride, dtype=u8, device=dev)
        self.anchor_fixed = None
        self.U = torch.zeros(n, dtype=i32, device=dev)
        self.idx = torch.zeros(S, n, dtype=i32, device=dev)
        self.old_aa = torch.zeros(S, n, dtype=u8, device=dev)
        self.lqf = torch.zeros(S, n, dtype=f32, device=dev)
        self.lqr = torch.zeros(S, n, dtype=f32, device=dev)
        self.log_acc = torch.zeros(n, dtype=f32, device=dev)
        self.accept = torch.zeros(n, dtype=u8, device=dev)
        T = self.T
        self.E_hist = torch.zeros(T + 1, n, dtype=f32, device=dev) if T is not None else None
        self.fit_hist = torch.zeros(T + 1, n, dtype=f32, device=dev) if T is not None else None
        self.best_E = torch.zeros(n, dtype=f32, device=dev)
        self.best_fit = torch.zeros(n, dtype=f32, device=dev)
        self.best_aa = torch.zeros(n, m.aa_stride, dtype=u8, device=dev)
        local_traj = (T is not None and 0 <= self.traj_chain < n)
        self.traj_aa = torch.zeros(T + 1, m.aa_stride, dtype=u8, device=dev) if local_traj else None
        self.t_dev = torch.zeros(1, dtype=i32, device=dev)
        # incremental CNN forward: per-row block keys of the max-pool and relu-mask rows, indexed like G / Gp
        self.inc = bool(m.cnn_inc)
        self.bkey = self.r1pool = self.dmask = self.mkpool = self.btab = None
        self.delta = bool(m.cnn_bwd_delta)
        # fused gradient combine: pas_reverse_accept assembles the proposal's row from the delta backward's sparse output
        # (no cnn_grad_combine_sparse_kernel launch).  PPDE_FUSE_COMBINE=0 keeps the separate kernel (A/B).
        # fused Potts field update: pas_propose does the work of ppde_potts_incremental (PPDE_FUSE_POTTS=0: separate kernel)
        self.fuse_potts = (m.has_potts and os.environ.get("PPDE_FUSE_POTTS", "1") != "0"
                           and m.L <= int(m.lib.ppde_pas_reverse_fuse_max_len()))
        compact = m.tune is None or m.tune.delta_layout == 0
        self.fuse_combine = (self.delta and compact and os.environ.get("PPDE_FUSE_COMBINE", "1") != "0"
                             and m.L <= int(m.lib.ppde_pas_reverse_fuse_max_len()) and m.n_nets <= 3)
        if self.inc:
            # per (row, net, channel): the two largest raw block keys (winner, runner-up) - ppde_cnn_forward_inc
            self.mkpool = torch.empty(rows * m.n_nets * 2 * m.C * 2, dtype=torch.int64, device=dev)
            # block table: btab[r][q] = the pool row whose slot holds block q (keys and relu-mask rows) of row r
            self.btab = torch.arange(rows, dtype=i32, device=dev).repeat_interleave(m.NB).contiguous()
            self.bkey = torch.empty(rows * m.n_nets * m.NB * 2 * m.C, dtype=torch.int64, device=dev)
            self.r1pool = torch.empty(rows * m.n_nets * m.P * 32, dtype=u8, device=dev)
            self.dmask = torch.zeros(n, dtype=i32, device=dev)
        self._allocated = True

    def _struct(self):
        m, n = self.m, self.n
        c = ChainsT(n=n, chain_offset=self.chain_offset, L=m.L, aa_stride=m.aa_stride)
        for name in ("aa", "aa_y", "row_cur", "G", "Gp", "E", "fit", "E_y", "fit_y", "Epotts_y", "E_fixed",
                     "fit_fixed", "aa_fixed", "U", "idx", "old_aa", "lqf", "lqr", "log_acc", "accept",
                     "best_E", "best_fit", "best_aa"):
            setattr(c, name, getattr(self, name).data_ptr())
        c.anchor_fixed = self.anchor_fixed.data_ptr() if self.anchor_fixed is not None else None
        c.E_hist = self.E_hist.data_ptr() if self.E_hist is not None else None
        c.fit_hist = self.fit_hist.data_ptr() if self.fit_hist is not None else None
        c.traj_aa = self.traj_aa.data_ptr() if self.traj_aa is not None else None
        c.traj_chain = self.traj_chain if self.traj_aa is not None else -1
        c.n_fixed = self.n_fixed
        c.row_wt = 2 * n
        return c

    # -- t = 0 --------------------------------------------------------------------------------------
    def init_population(self, aa0, anchor=None):
        """aa0: uint8 device tensor [n, aa_stride] (initial population; ppde.py:32-47).
        anchor: paper-mode `x` (ppde.py:35,76-77) — defaults to aa0, as in the reference."""
        m, n = self.m, self.n
        with torch.cuda.device(m.device):
            wt_row = m.wt[None, :m.L]
            all_wt = bool((aa0[:, :m.L] == wt_row).all().item())
            anchor = aa0 if anchor is None else anchor
            per_chain_anchor = self.paper and not bool((anchor[:, :m.L] == wt_row).all().item())
            self._alloc(1 + (n if per_chain_anchor else 0))
            self.aa.copy_(aa0)
            self.aa_fixed[0].copy_(m.wt)
            if per_chain_anchor:
                self.aa_fixed[1:].copy_(anchor)
                self.anchor_fixed = torch.arange(1, n + 1, dtype=torch.int32, device=m.device)
            ep = torch.empty(self.n_fixed, dtype=torch.float32, device=m.device)
            m.evaluate_into(self.aa_fixed, self.n_fixed, self.G, 2 * n, self.Gp, 2 * n, self.E_fixed, self.fit_fixed, ep,
                            bkey=self.bkey, r1pool=self.r1pool, mkpool=self.mkpool, btab=self.btab, ws=self.ws)
            if all_wt:
                self.row_cur.fill_(2 * n)
                self.E.copy_(self.E_fixed[0].expand(n)); self.fit.copy_(self.fit_fixed[0].expand(n))
            else:
                epn = torch.empty(n, dtype=torch.float32, device=m.device)
                m.evaluate_into(self.aa, n, self.G, 0, self.Gp, 0, self.E, self.fit, epn, bkey=self.bkey, r1pool=self.r1pool,
                                mkpool=self.mkpool, btab=self.btab, ws=self.ws)
                self.row_cur.copy_(torch.arange(n, dtype=torch.int32, device=m.device))
            self.best_E.copy_(self.E); self.best_fit.copy_(self.fit); self.best_aa.copy_(self.aa)
            if self.E_hist is not None:
                self.E_hist[0].copy_(self.E); self.fit_hist[0].copy_(self.fit)
            if self.traj_aa is not None:
                self.traj_aa[0].copy_(self.aa[self.traj_chain])
            self.t = 0
            self.t_dev.zero_()
            self.chains = self._struct()
            self._graph = None

    # -- one iteration ------------------------------------------------------------------------------
    def _params(self, t, uniforms=None, use_t_dev=False, full=True):
        """full=False (an iteration with the delta backward) + fuse_combine: the reverse kernel also assembles the gradient row."""
        p = PasParamsT(S=self.S, nmut_threshold=self.thr, paper_results=int(self.paper), t=int(t),
                       min_pos=self.min_pos, max_pos=self.max_pos, seed=self.seed,
                       uniforms=uniforms.data_ptr() if uniforms is not None else None,
                       t_dev=self.t_dev.data_ptr() if use_t_dev else None, full_trace=int(self.full_trace))
        p.fuse_potts = int(self.fuse_potts)
        if self.fuse_combine and not full:
            m = self.m
            vcap, rec, off = C.c_int32(0), C.c_int32(0), C.c_int64(0)
            _lib.check(self.lib.ppde_cnn_backward_delta_layout(C.byref(m.cnn), self.n, C.byref(vcap), C.byref(rec), C.byref(off)),
                       "cnn_backward_delta_layout")
            base = self.ws.grad_scratch(self.n).data_ptr()
            p.comb_nets, p.comb_vals, p.comb_wl = m.n_nets, base, base + 4 * off.value
            p.comb_vcap, p.comb_rec, p.comb_scale = vcap.value, rec.value, float(m.lamda) / m.n_nets
        return p

    def _launch_step(self, p, full=True):
        m, lib, c, n = self.m, self.lib, self.chains, self.n
        st = _stream()
        _lib.check(lib.ppde_pas_propose(C.byref(m.potts), C.byref(c), C.byref(p), st), "pas_propose")
        if m.has_potts and not p.fuse_potts:
            _lib.check(lib.ppde_potts_incremental(C.byref(m.potts), C.byref(c), C.byref(p), st), "potts_incremental")
        _lib.check(lib.ppde_step_rows(C.byref(c), _ptr(self.rows_y), st), "step_rows")
        self.cnn_forward_y(st)
        self.cnn_backward_y(st, full=full)
        _lib.check(lib.ppde_pas_reverse_accept(C.byref(m.potts), C.byref(c), C.byref(p), st), "pas_reverse_accept")

    def cnn_forward_y(self, st, dirty=True, parts=7):
        """CNN ensemble at the proposals aa_y: max-pool winners -> mkey (only the dirty blocks on the incremental path).
        dirty / parts select sub-kernels for per-kernel timing (bench.py); the defaults run everything."""
        m, n = self.m, self.n
        mk = self.ws.mkey(n)
        if self.inc:
            if dirty:
                _lib.check(self.lib.ppde_cnn_dirty(C.byref(m.cnn), _ptr(self.aa), _ptr(self.aa_y), m.aa_stride, n,
                                                   _ptr(self.dmask), st), "cnn_dirty")
            if parts:
                m.cnn_forward_pool(self.aa_y, n, mk, self.bkey, self.r1pool, self.dmask, self.row_cur, self.rows_y, 0, st,
                                   mkpool=self.mkpool, btab=self.btab, ws=self.ws, parts=0 if parts == 7 else parts)
        elif parts == 7:
            m.cnn_forward(self.aa_y, n, mk, st, ws=self.ws)

    def cnn_backward_y(self, st, do_fit=True, parts=7, full=True):
        """fit_y / E_y and the gradient rows of the proposals from the winners in mkey.  full=False (incremental path
        only): the rows are updated from the current state's rows by the delta backward."""
        m, n = self.m, self.n
        mk = self.ws.mkey(n)
        gp = _ptr(self.Gp) if m.has_potts else C.c_void_p(0)
        ep = _ptr(self.Epotts_y) if m.has_potts else C.c_void_p(0)
        if self.inc:
            pp = 0 if parts in (0, 7) else parts
            if self.delta and not full:
                if self.fuse_combine and parts in (4, 7):      # the combine runs inside pas_reverse_accept (_params(full=False))
                    pp = 3 if parts == 7 else 0
                    if parts == 4:
                        parts = 0
                m.cnn_backward_delta(self.aa, self.aa_y, n, mk, self.mkpool, gp, ep, _ptr(self.G), self.row_cur, self.rows_y,
                                     self.E_y, self.fit_y, self.r1pool, st, do_fit=do_fit, do_grad=bool(parts), btab=self.btab,
                                     ws=self.ws, parts=pp)
            else:
                m.cnn_backward_pool(self.aa_y, n, mk, gp, _ptr(self.rows_y), ep, _ptr(self.G), _ptr(self.rows_y),
                                    self.E_y, self.fit_y, self.r1pool, self.rows_y, 0, st, do_fit=do_fit, do_grad=bool(parts),
                                    btab=self.btab, ws=self.ws, parts=pp)
        elif do_fit and parts == 7:
            m.cnn_backward_combine(self.aa_y, n, mk, gp, _ptr(self.rows_y), ep, _ptr(self.G), _ptr(self.rows_y),
                                   self.E_y, self.fit_y, st, ws=self.ws)

    def full_backward_at(self, t):
        """Iteration t runs the exact backward (always without the delta path; every bwd_refresh-th iteration with it)."""
        return (not self.delta) or (t % self.m.bwd_refresh == self.m.bwd_refresh - 1)

    def _check_room(self, k):
        """The history rows E_hist / fit_hist / traj_aa [T+1, ...] are written at t+1 without a bound in the kernel."""
        if self.T is not None and self.t + k > self.T:
            raise ValueError(f"engine was allocated for num_steps={self.T}: cannot run iteration {self.t + k - 1} "
                             "(its history rows would fall outside the buffers)")

    def step(self, uniforms=None):
        """One MCMC iteration (eager launches). `uniforms`: optional float32 device tensor
        [S, n, 20L] replacing the in-kernel Philox proposal stream (parity mode)."""
        self._check_room(1)
        with torch.cuda.device(self.m.device):
            full = self.full_backward_at(self.t)
            self._launch_step(self._params(self.t, uniforms, full=full), full=full)
        self.t += 1

    def run_steps(self, k, use_graph=True):
        """k iterations; the fixed-S launch sequence is captured once in a CUDA graph and replayed,
        with the iteration counter living on the device."""
        self._check_room(k)
        if not use_graph:
            for _ in range(k):
                self.step()
            return
        with torch.cuda.device(self.m.device):
            self._graph_setup()
            self.t_dev.fill_(self.t)
            for i in range(k):
                full = self.full_backward_at(self.t + i)
                if full not in self._graph:                    # captured once per variant (exact / delta backward)
                    self._capture(full)
                self._graph[full].replay()
        self.t += k

    def _graph_setup(self):
        if self._graph is None:
            self.ws.mkey(self.n)
            self.ws.grad_scratch(self.n)
            if not self.inc:
                self.ws.r1mask(self.n)
            else:
                self.ws.inc_ws(self.n)
            self._graph_params = {full: self._params(0, None, use_t_dev=True, full=full) for full in (True, False)}
            self._graph = {}

    def _capture(self, full):
        """Record one iteration (exact or delta backward) into a CUDA graph; nothing is executed."""
        # capture_begin / capture_end directly: the `torch.cuda.graph` context manager also runs gc.collect(), a device
        # synchronize and torch.cuda.empty_cache() on entry (measured: 3 .. 650 ms of host time when another engine's pools
        # sit in the allocator cache); nothing is allocated while these launches are recorded, every buffer exists already
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            g.capture_begin()
            try:
                self._launch_step(self._graph_params[full], full=full)
                _lib.check(self.lib.ppde_counter_add(_ptr(self.t_dev), 1, _stream()), "counter_add")
            finally:
                g.capture_end()
        cur.wait_stream(side)
        self._graph[full] = g

    def prepare_graphs(self):
        """Capture every graph variant run_steps can need (delta and exact backward), so that no capture falls into a timed
        region later."""
        with torch.cuda.device(self.m.device):
            self._graph_setup()
            for full in ({True, False} if self.delta else {True}):
                if full not in self._graph:
                    self._capture(full)

    # -- results ------------------------------------------------------------------------------------
    def population_metrics(self):
        """(edit distance int32 [n], sequence hash int64 [n]) of the current states (device tensors)."""
        m = self.m
        dist = torch.empty(self.n, dtype=torch.int32, device=m.device)
        h = torch.empty(self.n, dtype=torch.int64, device=m.device)
        _lib.check(self.lib.ppde_population_metrics(_ptr(self.aa), m.aa_stride, self.n, m.L, _ptr(m.wt), _ptr(dist),
                                                    _ptr(h), _stream()), "population_metrics")
        return dist, h
