#!/usr/bin/env python
"""Long-run check of the cached state: T graph-replayed iterations (many exact-refresh periods) on a large population, then the
cached energies / Potts fields / gradient rows of a sample of chains against a from-scratch evaluation of their final states.
usage (GPU box): python tools/soak.py [chains=16384] [iterations=320] [sample=1024]   -> one JSON line"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppde_b200.engine import ChainEngine, PoEModel
from ppde_b200.synthetic import synthetic_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 320
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
L = 238
pr = synthetic_problem(L, seed=0)
m = PoEModel(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], 15.0, device="cuda:0")
eng = ChainEngine(m, n, 2, 0, False, seed=0, num_steps=T)
pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = pr["wt"]
eng.init_population(torch.from_numpy(pad).to(m.device))
eng.prepare_graphs()
torch.cuda.synchronize(); t0 = time.time()
eng.run_steps(T, use_graph=True)
torch.cuda.synchronize(); dt = time.time() - t0
sel = torch.linspace(0, n - 1, ns, device=m.device).long()
aa = eng.aa[sel].contiguous()
E, fit, G, Ep = m.energy(aa)
rc = eng.row_cur.long()[sel]
Gc = eng.G[rc].view(ns, L, 20)
gs = G.abs().amax(dim=(1, 2), keepdim=True)
Gpc = eng.Gp[rc]
Ef, ff, _, _ = E, fit, None, None
rel = lambda a, b, fl: float(((a - b).abs() / torch.clamp(b.abs(), min=fl)).max())
dist = (eng.aa[:, :L] != m.wt[None, :L]).sum(dim=1).float()
out = {"chains": n, "iterations": T, "exact_refresh_every": m.bwd_refresh, "sample": ns, "wall_s": round(dt, 3),
       "chain_steps_per_s": n * T / dt,
       "accept_rate_last": float(eng.accept.float().mean()), "mean_edit_distance": float(dist.mean()),
       "finite": bool(torch.isfinite(eng.E).all() and torch.isfinite(eng.G[eng.row_cur.long()]).all()),
       "cached_E_vs_fresh_rel": rel(eng.E[sel], E, abs(m.wt_H)), "cached_fit_vs_fresh_abs": float((eng.fit[sel] - fit).abs().max()),
       "cached_G_vs_fresh_rel_to_row_max": float(((Gc - G).abs() / gs).max()),
       "best_E_max": float(eng.best_E.max())}
print(json.dumps(out))
