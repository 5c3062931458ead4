#!/usr/bin/env python
"""Role-level cycle counters of the 2-CTA forward kernel (instrumented build, PROF=true template).

usage (GPU box): python tools/prof_fwd.py [chains] [L]
Prints, per role, the mean cycles per tile spent waiting on each barrier vs working."""
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
grid = 148
buf = torch.zeros(grid * 16, dtype=torch.int64, device=m.device)
tune = TuneT()
def run():
    _lib.check(lib.ppde_cnn_forward_tc(C.byref(m.cnn), _ptr(aad), m.aa_stride, n, _ptr(mk), _ptr(rm), C.byref(tune), _stream()), "fwd")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"plain kernel: {e0.elapsed_time(e1):.3f} ms for {n} chains")
tune.prof = buf.data_ptr()
run(); torch.cuda.synchronize()
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"instrumented: {e0.elapsed_time(e1):.3f} ms")
tune.prof = None
c = buf.cpu().numpy().reshape(grid, 16)[:144]
P = L - 4; tpc = (P + 127) // 128
tiles = n / 12 * tpc        # per cluster (12 clusters per combo at L=238)
lead, foll = c[0::2], c[1::2]
def pt(x): return f"{x.mean() / tiles:8.0f}"
print("cycles per tile (mean over CTAs)")
print(" leader MMA thread: wait dempty", pt(lead[:, 3]), " wait fullL", pt(lead[:, 4]), " wait fullR", pt(lead[:, 5]), " issue+commit", pt(lead[:, 6]), " total", pt(lead[:, 7]))
print(" forwarder (rank1): wait fullL", pt(foll[:, 3]), " remote arrive", pt(foll[:, 4]))
for nm, cc in (("leader", lead), ("rank1 ", foll)):
    print(f" {nm} epilogue w0: wait dfull", pt(cc[:, 0]), " ld+scan", pt(cc[:, 1]), " arrive", pt(cc[:, 2]))
    print(f" {nm} producer w0 : wait empty", pt(cc[:, 8]), " produce", pt(cc[:, 9]), " fence+arrive", pt(cc[:, 10]))
    print(f" {nm} producer w15: wait empty", pt(cc[:, 11]), " produce", pt(cc[:, 12]), " fence+arrive", pt(cc[:, 13]))
