#!/usr/bin/env python
"""Incremental CNN forward inside a realistic PPDE step: time of the whole forward (dirty + scan + tensor-core + merge
kernels) on fresh proposals, and role-level cycle counters of the tensor-core kernel (runtime-instrumented through
ppde_tune_t.prof).   usage (GPU box): python tools/prof_inc.py [chains]   (PPDE_INC_DEBUG=1|2|4: experiments)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppde_b200 import _lib
from ppde_b200._lib import TuneT
from ppde_b200.engine import ChainEngine, PoEModel, _ptr, _stream
from ppde_b200.synthetic import synthetic_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
L = 238
pr = synthetic_problem(L, seed=0)
m = PoEModel(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], 15.0, device="cuda:0")
lib = m.lib
eng = ChainEngine(m, n, 2, 0, False, seed=0)
pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = pr["wt"]
dbg = os.environ.pop("PPDE_INC_DEBUG", None)           # warm up to a realistic population with correct kernels
eng.init_population(torch.from_numpy(pad).to(m.device))
eng.run_steps(8, use_graph=False)
torch.cuda.synchronize()
st = _stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def step(timed_forward):
    """one eager iteration; returns the forward's milliseconds"""
    p = eng._params(eng.t)
    _lib.check(lib.ppde_pas_propose(C.byref(m.potts), C.byref(eng.chains), C.byref(p), st), "propose")
    _lib.check(lib.ppde_potts_incremental(C.byref(m.potts), C.byref(eng.chains), C.byref(p), st), "inc")
    _lib.check(lib.ppde_step_rows(C.byref(eng.chains), _ptr(eng.rows_y), st), "rows")
    e0.record(); timed_forward(); e1.record()
    eng.cnn_backward_y(st)
    _lib.check(lib.ppde_pas_reverse_accept(C.byref(m.potts), C.byref(eng.chains), C.byref(p), st), "rev")
    torch.cuda.synchronize()
    eng.t += 1
    return e0.elapsed_time(e1)

if dbg: m.tune = TuneT(dbg=int(dbg))
ts = [step(lambda: eng.cnn_forward_y(st)) for _ in range(5)]
nd = torch.tensor([bin(int(v) & 0xFFFFFFFF).count("1") for v in eng.dmask.cpu().numpy()[:4096]]).float()
print(f"chains {n} dbg {dbg}: incremental forward {min(ts):.3f} ms (median {sorted(ts)[2]:.3f}); dirty blocks / chain mean {nd.mean():.2f} max {nd.max():.0f}")
if dbg:
    sys.exit(0)
buf = torch.zeros(148 * 16, dtype=torch.int64, device=m.device)
def fwd_prof():
    m.tune = TuneT(prof=buf.data_ptr())          # only around the forward: the backward kernels read the same ppde_tune_t
    eng.cnn_forward_y(st)
    m.tune = None
t_inst = step(fwd_prof)
c = buf.cpu().numpy().reshape(148, 16)[:144]
lead, peer = c[0::2], c[1::2]
chains, tiles = lead[:, 10].mean(), lead[:, 9].mean()
print(f"instrumented: {t_inst:.3f} ms; per cluster: {chains:.0f} chains, {tiles:.0f} tiles; tensor-core kernel main loop {lead[:, 8].mean():.0f} cycles")
def pt(x): return f"{x.mean() / tiles:7.0f}"
print("cycles per TILE of 8 blocks (mean over clusters)")
print(" epilogue (leader w0): prefetch", pt(lead[:, 0]), " wait dfull", pt(lead[:, 1]), " blocks+arrive", pt(lead[:, 2]))
print(" epilogue (peer   w0): prefetch", pt(peer[:, 0]), " wait dfull", pt(peer[:, 1]), " blocks+arrive", pt(peer[:, 2]))
print(" MMA thread          : wait dempty", pt(lead[:, 4]), " wait fullL", pt(lead[:, 5]), " wait fullR", pt(lead[:, 6]), " issue+commit", pt(lead[:, 7]))
print(" producer w0 (leader): wait empty", pt(lead[:, 12]), " total", pt(lead[:, 13]))
