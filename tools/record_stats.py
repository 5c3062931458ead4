#!/usr/bin/env python
"""Statistics of the compact delta-backward records in steady state: touched positions, entries, and how many entries are the
same (position, channel) on both sides (a winner that stays on a conv row whose relu mask changed).  usage: python tools/record_stats.py [chains] [steps]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppde_b200 import _lib
from ppde_b200.engine import ChainEngine, PoEModel
from ppde_b200.synthetic import synthetic_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 40
L = 238
pr = synthetic_problem(L, seed=0)
m = PoEModel(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], 15.0, device="cuda:0")
eng = ChainEngine(m, n, 2, 0, False, seed=0)
pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = pr["wt"]
eng.init_population(torch.from_numpy(pad).to(m.device))
eng.run_steps(T, use_graph=False)
torch.cuda.synchronize()
vcap, rec, off = C.c_int32(0), C.c_int32(0), C.c_int64(0)
_lib.check(m.lib.ppde_cnn_backward_delta_layout(C.byref(m.cnn), n, C.byref(vcap), C.byref(rec), C.byref(off)), "layout")
sc = eng.ws.grad_scratch(n)
wl = sc[off.value:].view(torch.int16)[: n * m.n_nets * rec.value].cpu().numpy().view(np.uint16).reshape(n * m.n_nets, rec.value)
npos, nent, dup, d0ent, tiles = [], [], [], [], []
NW, NT = 6, 48                      # tc::BD_NW, tc::BD_NT: record = npos | pos[npos] | woff[NW ntile + 1] | list[nent] | ...
for r in wl[:3000]:
    p = int(r[0]); nt = (p + NT - 1) // NT; nwo = NW * nt + 1
    woff = r[1 + p: 1 + p + nwo].astype(int); e = int(woff[-1])
    lst = r[1 + p + nwo: 1 + p + nwo + e].astype(int)
    nd = 0
    for g in range(nwo - 1):           # entries of (tile, producer warp) group g, slot in bits 12-14, side in bit 15, channel in bits 0-8
        seg = lst[woff[g]: woff[g + 1]]
        for sl in range(8):
            ss = seg[((seg >> 12) & 7) == sl]
            ch_y = set(ss[(ss & 0x8000) == 0] & 0x1FF); ch_x = set(ss[(ss & 0x8000) != 0] & 0x1FF)
            nd += len(ch_y & ch_x)
    npos.append(p); nent.append(e); dup.append(nd); tiles.append(nt)
npos, nent, dup, tiles = map(np.array, (npos, nent, dup, tiles))
print(f"records {len(npos)}: touched positions mean {npos.mean():.1f} (p90 {np.percentile(npos, 90):.0f}, max {npos.max()}), entries mean {nent.mean():.1f} "
      f"(p90 {np.percentile(nent, 90):.0f}), same (position, channel) on both sides: {dup.mean():.1f} pairs = {2 * dup.sum() / max(nent.sum(), 1) * 100:.1f} % of the entries; "
      f"tiles mean {tiles.mean():.2f}, > 1 tile {np.mean(tiles > 1) * 100:.1f} %, empty {np.mean(npos == 0) * 100:.1f} %")
