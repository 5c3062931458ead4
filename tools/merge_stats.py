#!/usr/bin/env python
"""Rescan rate of the top-2 merge in steady state (counters through ppde_tune_t.prof).  usage: python tools/merge_stats.py [chains] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppde_b200._lib import TuneT
from ppde_b200.engine import ChainEngine, PoEModel, _stream
from ppde_b200.synthetic import synthetic_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
L = 238
pr = synthetic_problem(L, seed=0)
m = PoEModel(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], 15.0, device="cuda:0")
eng = ChainEngine(m, n, 2, 0, False, seed=0)
pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = pr["wt"]
eng.init_population(torch.from_numpy(pad).to(m.device))
buf = torch.zeros(16, dtype=torch.int64, device=m.device)
chan = n * m.n_nets * 2 * m.C
for t in range(T):
    if t % 10 == 9:
        buf.zero_()
        m.tune = TuneT(prof=buf.data_ptr())
    eng.run_steps(1, use_graph=False)
    if m.tune is not None:
        torch.cuda.synchronize()
        m.tune = None
        c = buf.cpu().numpy()
        nd = np.array([bin(int(v) & 0xFFFFFFFF).count("1") for v in eng.dmask.cpu().numpy()[:4096]])
        print(f"t={t}: rescans {c[0] / chan * 100:.2f} % of the channels, "
              f"dirty blocks / chain {nd.mean():.2f} (max {nd.max()}, > 4: {(nd > 4).mean() * 100:.1f} %, > 6: {(nd > 6).mean() * 100:.1f} %)")
