#!/usr/bin/env python
"""Full Potts re-evaluation: dense tcgen05 GEMM vs row-gather kernel (BASELINE.json configs[4]-style sweep).
usage (GPU box): python tools/bench_potts_full.py [L ...]   (Potts-only: lamda = 0, synthetic couplings)"""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppde_b200.engine import PoEModel, _ptr, _stream
from ppde_b200.synthetic import synthetic_problem

Ls = [int(a) for a in sys.argv[1:]] or [64, 128, 238, 512]
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
for L in Ls:
    pr = synthetic_problem(L, seed=0)
    m = PoEModel(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], 0.0, device="cuda:0")
    for n in ([1024, 16384, 65536] if L <= 512 else [1024, 16384]):
        rng = np.random.default_rng(0)
        aa = np.tile(pr["wt"], (n, 1)).astype(np.uint8)
        idx = rng.integers(0, L, size=(n, 8)); val = rng.integers(0, 20, size=(n, 8))
        np.put_along_axis(aa, idx, val.astype(np.uint8), axis=1)
        pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = aa
        aad = torch.from_numpy(pad).to(m.device)
        Gp = torch.empty(n, m.D, dtype=torch.float32, device=m.device)
        Ep = torch.empty(n, dtype=torch.float32, device=m.device)
        res = {}
        for impl in ("gather", "dense"):
            for _ in range(2): m.potts_full(aad, n, _ptr(Gp), _ptr(Ep), _stream(), impl=impl)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps): m.potts_full(aad, n, _ptr(Gp), _ptr(Ep), _stream(), impl=impl)
            e1.record(); torch.cuda.synchronize()
            res[impl] = e0.elapsed_time(e1) / reps
        flops = 2.0 * n * m.D * m.D
        tf = flops / (res["dense"] * 1e-3) / 1e12
        print(json.dumps({"L": L, "D": m.D, "chains": n, "gather_ms": round(res["gather"], 3), "dense_ms": round(res["dense"], 3),
                          "dense_alg_TFLOPs": round(tf, 1), "frac_of_measured_bf16_peak": round(tf / peaks["bf16_tflops"], 3),
                          "note": "2 fp16 passes per algorithmic flop (hi/lo split): executed = 2x"}))
    del m
    torch.cuda.empty_cache()
