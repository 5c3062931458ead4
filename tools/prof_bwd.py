#!/usr/bin/env python
"""Role-level cycle counters of the tensor-core backward kernel (instrumented build).
usage (GPU box): python tools/prof_bwd.py [chains] [L]"""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppde_b200 import _lib
from ppde_b200._lib import TuneT
from ppde_b200.engine import PoEModel, _ptr, _stream
from ppde_b200.synthetic import synthetic_problem

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = int(sys.argv[2]) if len(sys.argv) > 2 else 238
pr = synthetic_problem(L, seed=0)
m = PoEModel(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], 15.0, device="cuda:0")
lib = m.lib
rng = np.random.default_rng(0)
aa = np.tile(pr["wt"], (n, 1)).astype(np.uint8)
for b in range(n):
    pos = rng.integers(0, L, size=10); aa[b, pos] = rng.integers(0, 20, size=10)
pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = aa
aad = torch.from_numpy(pad).to(m.device)
mk = m.ws.mkey(n); rm = m.ws.r1mask(n)
G = torch.empty(n, m.NE, dtype=torch.float32, device=m.device)
Gp = torch.zeros(n, m.D, dtype=torch.float32, device=m.device)
E = torch.empty(n, dtype=torch.float32, device=m.device); fit = torch.empty_like(E); Ep = torch.zeros_like(E)
st = _stream()
def fwd():
    _lib.check(lib.ppde_cnn_forward_tc(C.byref(m.cnn), _ptr(aad), m.aa_stride, n, _ptr(mk), _ptr(rm), None, st), "fwd")
def bwd():
    m.cnn_backward_combine(aad, n, mk, _ptr(Gp), C.c_void_p(0), _ptr(Ep), _ptr(G), C.c_void_p(0), E, fit, st)
fwd()
for _ in range(3): bwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); bwd(); e1.record(); torch.cuda.synchronize()
print(f"plain backward (fit + sort + tc + combine): {e0.elapsed_time(e1):.3f} ms for {n} chains")
grid = 148
buf = torch.zeros(grid * 16, dtype=torch.int64, device=m.device)
m.tune = TuneT(prof=buf.data_ptr())
bwd(); torch.cuda.synchronize()
e0.record(); bwd(); e1.record(); torch.cuda.synchronize()
print(f"instrumented: {e0.elapsed_time(e1):.3f} ms")
m.tune = None
c = buf.cpu().numpy().reshape(grid, 16)[:147]
P = L - 4; tpc = (P + 63) // 64
tiles = n / 49 * tpc
def pt(x): return f"{x.mean() / tiles:8.0f}"
print("cycles per 64-position tile (mean over CTAs); tiles per CTA =", tiles)
print(" epilogue t0 : wait dfull", pt(c[:, 0]), " tmem ld + sY", pt(c[:, 1]), " col2im", pt(c[:, 2]), " flush (per tile avg)", pt(c[:, 3]))
print(" MMA thread  : wait dempty", pt(c[:, 4]), " wait full", pt(c[:, 5]), " issue+commit", pt(c[:, 6]))
print(" producer w0 : wait empty", pt(c[:, 8]), " gather", pt(c[:, 9]), " fence+arrive", pt(c[:, 10]), " cp.async wait + bars", pt(c[:, 11]))
print(" producer w15: wait empty", pt(c[:, 12]), " gather", pt(c[:, 13]), " fence+arrive", pt(c[:, 14]), " cp.async wait + bars", pt(c[:, 15]))
