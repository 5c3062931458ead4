#!/usr/bin/env python
"""Repeats the e2e scenario (fresh engine from the wild type, K graph-replayed iterations) and prints per-iteration GPU times
(CUDA events) and host wall times, to separate device-side stalls from host-side hiccups.
usage (GPU box): python tools/e2e_jitter.py [workload] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ppde_b200.engine import ChainEngine
from ppde_b200.energy import ProteinProductOfExperts
wlname = sys.argv[1] if len(sys.argv) > 1 else "ube4b_potts_poe_4k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
wl = dict(bench.WORKLOADS[wlname]); pr = bench.build_problem(wl)
dev = torch.device("cuda", 0)
energy = ProteinProductOfExperts.from_arrays(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], wl["lamda"], device=dev)
m = energy.model; n, L, K = wl["chains"], wl["L"], 10
pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = pr["wt"]
aa0 = torch.from_numpy(pad).to(dev)
for rep in range(reps):
    t0 = time.perf_counter()
    eng = ChainEngine(m, n, wl["pas"], wl["nmut"], wl["paper"], seed=rep, num_steps=K)
    eng.init_population(aa0)
    t1 = time.perf_counter()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    evs[0].record()
    for i in range(K):
        eng.run_steps(1, use_graph=True)
        evs[i + 1].record()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(K)]
    flag = "  <-- SLOW" if (t3 - t0) > 0.06 else ""
    print(f"rep {rep:2d}: init {1e3*(t1-t0):6.1f} ms  enqueue {1e3*(t2-t1):6.1f} ms  drain {1e3*(t3-t2):7.1f} ms  gpu per step: "
          + " ".join(f"{x:.2f}" for x in ms) + flag, flush=True)
    del eng
