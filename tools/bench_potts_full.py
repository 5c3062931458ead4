#!/usr/bin/env python
"""Potts expert alone (BASELINE.json configs[4]-style sweep): full re-evaluation as a dense tcgen05 GEMM vs the row-gather
kernel, and the sampler's incremental field update (k <= S gathered J row differences per chain) for the same chains.
usage (GPU box): python tools/bench_potts_full.py [L ...]   (Potts-only: lamda = 0, synthetic couplings)
Under torchrun (N ranks, one per GPU) every rank runs the same sweep on its own chains (weak scaling: chains independent, couplings
replicated, no collective in the path); times are the maximum over the ranks, `chains` is per GPU."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppde_b200 import _lib
from ppde_b200.engine import ChainEngine, PoEModel, _ptr, _stream
from ppde_b200.synthetic import synthetic_problem

Ls = [int(a) for a in sys.argv[1:]] or [64, 128, 238, 512]
WS, RANK, LOCAL = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
DEV = f"cuda:{LOCAL}"
torch.cuda.set_device(LOCAL)
if WS > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device(DEV))
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
for L in Ls:
    pr = synthetic_problem(L, seed=0)
    m = PoEModel(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], 0.0, device=DEV)
    # (the engine allocates the CNN pools too: 262,144 chains fit up to L = 128, the max-pool cache in blocks of 8 needs 45 GB per 64k chains at L = 238)
    for n in ([1024, 16384, 65536, 262144] if L <= 128 else [1024, 16384, 65536, 131072] if L <= 238 else ([1024, 16384, 65536] if L <= 512 else [1024, 16384])):
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
        res["incremental"] = float("nan")
        if m.C <= 512:        # the engine evaluates the CNN of the wild type once; the fp32 SIMT CNN (C > 256) needs C*68*4 B of smem
            # the sampler's incremental path on the same population size: propose (pas=2, up to 3 moves) then update the field
            del Gp
            eng = ChainEngine(m, n, 2, 0, False, seed=0)
            wt = np.zeros((n, m.aa_stride), dtype=np.uint8); wt[:, :L] = pr["wt"]
            eng.init_population(torch.from_numpy(wt).to(m.device))
            st = _stream()
            inc = []
            for rep in range(4):
                p = eng._params(rep)
                _lib.check(m.lib.ppde_pas_propose(C.byref(m.potts), C.byref(eng.chains), C.byref(p), st), "propose")
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(m.lib.ppde_potts_incremental(C.byref(m.potts), C.byref(eng.chains), C.byref(p), st), "inc")
                e1.record(); torch.cuda.synchronize()
                inc.append(e0.elapsed_time(e1))
            res["incremental"] = min(inc[1:])
            del eng
        if WS > 1:                 # device times, maximum over the ranks
            tt = torch.tensor([res["gather"], res["dense"], res["incremental"] if res["incremental"] == res["incremental"] else -1.0], device=DEV)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            res["gather"], res["dense"] = float(tt[0]), float(tt[1])
            if float(tt[2]) >= 0: res["incremental"] = float(tt[2])
            if RANK: continue
        flops = 2.0 * n * m.D * m.D * 1.0
        tf = flops / (res["dense"] * 1e-3) / 1e12
        print(json.dumps({"L": L, "D": m.D, "n_gpus": WS, "chains": n, "chains_total": n * WS,
                          "dense_alg_TFLOPs_total": round(tf * WS, 1), "gather_ms": round(res["gather"], 3), "dense_ms": round(res["dense"], 3),
                          "incremental_ms": None if res["incremental"] != res["incremental"] else round(res["incremental"], 3),
                          "incremental_GBs": None if res["incremental"] != res["incremental"] else round((8 * m.D + L) * n / (res["incremental"] * 1e-3) / 1e9, 1),
                          "dense_alg_TFLOPs": round(tf, 1), "frac_of_measured_bf16_peak": round(tf / peaks["bf16_tflops"], 3),
                          "note": "2 fp16 passes per algorithmic flop (hi/lo split): executed = 2x"}))
    del m
    torch.cuda.empty_cache()
if WS > 1:
    dist.barrier()
    dist.destroy_process_group()
