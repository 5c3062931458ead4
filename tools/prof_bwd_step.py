#!/usr/bin/env python
"""Role-level cycle counters of the tensor-core backward kernel INSIDE a realistic PPDE step (delta or exact mode).
usage (GPU box): python tools/prof_bwd_step.py [chains] [full|delta]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppde_b200 import _lib
from ppde_b200._lib import TuneT
from ppde_b200.engine import ChainEngine, PoEModel, _ptr, _stream
from ppde_b200.synthetic import synthetic_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
full = (sys.argv[2] == "full") if len(sys.argv) > 2 else False
L = 238
pr = synthetic_problem(L, seed=0)
m = PoEModel(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], 15.0, device="cuda:0")
lib = m.lib
eng = ChainEngine(m, n, 2, 0, False, seed=0)
pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = pr["wt"]
eng.init_population(torch.from_numpy(pad).to(m.device))
eng.run_steps(8, use_graph=False)
torch.cuda.synchronize()
st = _stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def step():
    p = eng._params(eng.t, full=full)
    _lib.check(lib.ppde_pas_propose(C.byref(m.potts), C.byref(eng.chains), C.byref(p), st), "propose")
    _lib.check(lib.ppde_potts_incremental(C.byref(m.potts), C.byref(eng.chains), C.byref(p), st), "inc")
    _lib.check(lib.ppde_step_rows(C.byref(eng.chains), _ptr(eng.rows_y), st), "rows")
    eng.cnn_forward_y(st)
    eng.cnn_backward_y(st, do_fit=True, parts=1, full=full)
    e0.record(); eng.cnn_backward_y(st, do_fit=False, parts=2, full=full); e1.record()
    eng.cnn_backward_y(st, do_fit=False, parts=4, full=full)
    _lib.check(lib.ppde_pas_reverse_accept(C.byref(m.potts), C.byref(eng.chains), C.byref(p), st), "rev")
    torch.cuda.synchronize()
    eng.t += 1
    return e0.elapsed_time(e1)
ts = [step() for _ in range(4)]
grid = 148
buf = torch.zeros(grid * 16, dtype=torch.int64, device=m.device)
m.tune = TuneT(prof=buf.data_ptr())
t_inst = step()
m.tune = None
c = buf.cpu().numpy().reshape(grid, 16)[:147]
P = L - 4; tpc = (P + 63) // 64
tiles = n / 49 * tpc
compact = (not full) and os.environ.get("PPDE_BWD_DELTA_COMPACT", "1") != "0" and eng.delta
if compact:
    tiles = float(c[:, 7].mean())          # the compact delta kernel counts its own tiles (normally one per chain and net)
print(f"backward tensor-core kernel ({'exact' if full else 'delta'}), {n} chains: {min(ts):.3f} ms; instrumented {t_inst:.3f} ms; tiles per CTA {tiles:.0f}")
def pt(x): return f"{x.mean() / tiles:8.0f}"
print(f"cycles per 64-column tile (mean over CTAs); chains per CTA {n / 49:.0f}")
print(" epilogue t0 : wait dfull", pt(c[:, 0]), " tmem ld", pt(c[:, 1]), " col2im", pt(c[:, 2]), " flush (per tile avg)", pt(c[:, 3]))
print(" MMA thread  : wait dempty", pt(c[:, 4]), " wait full", pt(c[:, 5]), " issue+commit", pt(c[:, 6]))
print(" producer w0 : wait empty", pt(c[:, 8]), " gather+rows", pt(c[:, 9]), " fence+arrive", pt(c[:, 10]), " wait record", pt(c[:, 11]))
print(" producer w15: wait empty", pt(c[:, 12]), " gather+rows", pt(c[:, 13]), " fence+arrive", pt(c[:, 14]), " wait record", pt(c[:, 15]))
