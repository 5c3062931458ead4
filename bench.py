#!/usr/bin/env python
"""Throughput of the PPDE product-of-experts MCMC step (BASELINE.json metric) on B200.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm (oracle port) on host cores

A "step" is one MCMC iteration (path-auxiliary proposal of S=2*pas-1 sub-steps, energy+gradient
at y for every chain, reverse proposal, MH accept, state commit) over all chains of the workload.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric/target is quoted on; fits one GPU.
    "gfp_potts_poe_64k": dict(L=238, chains=65536, lamda=15.0, pas=2, nmut=0, paper=False,
                              desc="GFP-length (L=238) Potts PoE, lambda=15, pas=2, 65536 chains/GPU, synthetic weights"),
    # BASELINE.json configs[1]
    "ube4b_potts_poe_4k": dict(L=104, chains=4096, lamda=0.5, pas=2, nmut=10, paper=False, window=(22, 97),
                               desc="UBE4B-length (L=104, window 23..98) Potts PoE, lambda=0.5, nmut=10, 4096 chains/GPU"),
    # BASELINE.json configs[0] (the reference's CPU-runnable README case)
    "pabp_readme_128": dict(L=96, chains=128, lamda=5.0, pas=2, nmut=0, paper=False,
                            desc="PABP-length (L=96) README command, lambda=5, 128 chains"),
    # BASELINE.json configs[3]
    "gfp_paper_pas10": dict(L=238, chains=16384, lamda=15.0, pas=10, nmut=0, paper=True,
                            desc="GFP-length Potts+CNN PoE, pas=10 (S=19), paper_results soft mode, 16384 chains/GPU"),
}
METRIC = "PPDE chain-steps/sec (Potts PoE)"
UNIT = "chain-steps/s"


def peaks():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_problem(wl):
    from ppde_b200.synthetic import synthetic_problem
    return synthetic_problem(wl["L"], seed=0, window=wl.get("window"))


def cpu_port_throughput(wl, pr, n_chains, steps, warmup):
    """The reference algorithm (oracle port, torch CPU ops, all host threads) on a bounded sample."""
    from oracle import ppde_port as port
    w = port.Weights(wt=pr["wt"], J=pr["J"], h=pr["h"], win_lo=pr["win_lo"], cnn=pr["cnn"], lamda=wl["lamda"])
    en = port.PortEnergy(w)
    smp = port.PortSampler(wl["pas"], wl["nmut"], wl["paper"], seed=0, fixed_S=False)
    x0 = en.wt_onehot.repeat(n_chains, 1, 1)
    cur = x0.clone()
    t = 0
    for _ in range(warmup):
        _, cur = smp.step(en, t, cur, x0); t += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        _, cur = smp.step(en, t, cur, x0); t += 1
    dt = time.perf_counter() - t0
    return n_chains * steps / dt, dt / steps * 1e3


def use_all_host_threads():
    """all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1 to its workers)"""
    try:
        torch.set_num_threads(max(torch.get_num_threads(), len(os.sched_getaffinity(0))))
    except Exception:
        pass
    return torch.get_num_threads()


def cpu_reference_throughput(wl, pr, n_chains, steps, warmup):
    """CPU arm: the UNMODIFIED reference (oracle/_ref, staged by oracle/stage_ref.py: PPDE_PAS.run + ProteinProductOfExperts,
    ppde/protein_samplers/ppde.py:24-192, ppde/energy.py:72-108) when it is there, else the oracle port.
    -> (chain-steps/s, ms per step, kind, what)"""
    from oracle import ref_arm
    if ref_arm.available() and os.environ.get("PPDE_BENCH_CPU", "reference") != "port":
        dt, _ = ref_arm.time_steps(pr, n_chains, wl["lamda"], wl["pas"], wl["nmut"], wl["paper"], steps, warmup)
        return n_chains * steps / dt, dt / steps * 1e3, "reference", \
            "unmodified reference PPDE_PAS.run + ProteinProductOfExperts (oracle/_ref), device cpu, steady-state iterations"
    v, ms = cpu_port_throughput(wl, pr, n_chains, steps, warmup)
    return v, ms, "port", "oracle port (torch CPU ops in the reference's order)"


def run_reference(args, wl, pr):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_cpu = args.cpu_chains
    cores = use_all_host_threads()
    val, ms, kind, what = cpu_reference_throughput(wl, pr, n_cpu, args.steps, args.warmup)
    sample = (f"{n_cpu} chains x {args.steps} iterations of the same workload (L={wl['L']}, pas={wl['pas']}, lambda={wl['lamda']}, "
              f"nmut={wl['nmut']}, same synthetic weights); the GPU arm runs {wl['chains']} chains per GPU: chain-steps/s is per chain, "
              f"the CPU is saturated at this batch; {what}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": args.workload, "desc": wl["desc"],
                                                             "cpu_sample_chains": n_cpu},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gfp_potts_poe_64k", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=None, help="chains per GPU (default: the workload's)")
    ap.add_argument("--cpu-chains", type=int, default=None,
                    help="bounded CPU sample size (default: min(1024, the workload's chains); >= 1024 saturates the host cores)")
    ap.add_argument("--cpu-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--strong", action="store_true", help="strong scaling: the workload's chains are the GLOBAL population, "
                                                          "sharded over the ranks (default: that many chains PER GPU, weak)")
    ap.add_argument("--global-population", action="store_true",
                    help="e2e leg: every rank passes the GLOBAL population and gets the global 6-tuple back (the reference's "
                         "calling convention; default: each rank passes its shard, ppde_local_population)")
    ap.add_argument("--log-every", type=int, default=0, help="e2e leg: run the log_every population report (device kernels + "
                                                             "NCCL all-gathers) every this many iterations (0 = never)")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.chains:
        wl["chains"] = args.chains
    if args.cpu_chains is None:
        args.cpu_chains = min(1024, wl["chains"])
    pr = build_problem(wl)
    if args.impl == "reference":
        run_reference(args, wl, pr)
        return

    import torch.distributed as dist
    from ppde_b200 import _lib
    from ppde_b200.engine import ChainEngine, PoEModel
    from ppde_b200.energy import ProteinProductOfExperts
    from ppde_b200.sampler import PPDE_PAS

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    dev = torch.device("cuda", local)
    lib = _lib.load()

    n = wl["chains"]                                       # per GPU (weak scaling: per-GPU work fixed)
    if args.strong:                                        # strong scaling: the workload's chains are the global population
        if n % world:
            raise SystemExit("--strong needs chains divisible by the number of GPUs")
        n //= world
    L = wl["L"]
    energy = ProteinProductOfExperts.from_arrays(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], wl["lamda"], device=dev)
    m = energy.model
    K, Wm = args.steps, args.warmup
    eng = ChainEngine(m, n, wl["pas"], wl["nmut"], wl["paper"], seed=0, chain_offset=rank * n, num_steps=None)
    pad = np.zeros((n, m.aa_stride), dtype=np.uint8)
    pad[:, :L] = pr["wt"]
    eng.init_population(torch.from_numpy(pad).to(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (inputs already in HBM) -------------------------------------
    c0 = lib.ppde_last_launch_count()
    eng.run_steps(max(Wm, 3), use_graph=True)              # warm-up (captures the CUDA graph once)
    eng.prepare_graphs()                                   # ... and the exact-refresh variant
    # The delta backward is refreshed by an exact backward every `bwd_refresh`-th iteration (t % R == R-1).  The timed K
    # iterations must carry their share of that work: the iteration counter is placed so that they contain exactly
    # ceil(K / R) refresh iterations (>= the long-run average of K / R; never fewer).
    refreshes = 0
    if eng.delta:
        R = m.bwd_refresh
        r_ = -(-K // R)
        eng.t = (eng.t // R + 1) * R + max(0, R * r_ - K)
        refreshes = sum(1 for i in range(K) if eng.full_backward_at(eng.t + i))
        assert refreshes == r_, (refreshes, r_)
    launches_per_step = None
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    eng.run_steps(K, use_graph=True)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = n * world * K / (ms * 1e-3)
    acc_rate = float(eng.accept.float().mean().item())
    mean_dist = float(eng.population_metrics()[0].float().mean().item())

    # ---- per-kernel durations, live (eager launches bracketed by CUDA events on the launch stream) ----
    breakdown, roof, refresh_ms = {}, None, None
    pk = peaks()
    if not args.no_breakdown:
        import ctypes as C
        from ppde_b200.engine import _ptr, _stream
        st = _stream()
        p_holder = {}
        seq = [("pas_propose", lambda: _lib.check(lib.ppde_pas_propose(C.byref(m.potts), C.byref(eng.chains), C.byref(p_holder["p"]), st), "propose")),
               # (with the fused Potts update the propose kernel has done this work already: only the row indices remain)
               ("potts_incremental", lambda: ((None if p_holder["p"].fuse_potts else
                                               _lib.check(lib.ppde_potts_incremental(C.byref(m.potts), C.byref(eng.chains), C.byref(p_holder["p"]), st), "inc")),
                                              _lib.check(lib.ppde_step_rows(C.byref(eng.chains), _ptr(eng.rows_y), st), "rows")))]
        if eng.inc:
            seq += [("cnn_dirty", lambda: eng.cnn_forward_y(st, dirty=True, parts=0)),
                    ("cnn_inc_scan", lambda: eng.cnn_forward_y(st, dirty=False, parts=1)),
                    ("cnn_forward_inc_tc", lambda: eng.cnn_forward_y(st, dirty=False, parts=2)),
                    ("cnn_inc_merge", lambda: eng.cnn_forward_y(st, dirty=False, parts=4)),
                    ("cnn_fit", lambda: eng.cnn_backward_y(st, do_fit=True, parts=0, full=not eng.delta)),
                    ("cnn_winner_sort", lambda: eng.cnn_backward_y(st, do_fit=False, parts=1, full=not eng.delta)),
                    ("cnn_backward_tc", lambda: eng.cnn_backward_y(st, do_fit=False, parts=2, full=not eng.delta)),
                    ("cnn_grad_combine", lambda: eng.cnn_backward_y(st, do_fit=False, parts=4, full=not eng.delta))]
        else:
            seq += [("cnn_forward", lambda: eng.cnn_forward_y(st)), ("cnn_backward_combine", lambda: eng.cnn_backward_y(st))]
        seq += [("pas_reverse_accept", lambda: _lib.check(lib.ppde_pas_reverse_accept(C.byref(m.potts), C.byref(eng.chains), C.byref(p_holder["p"]), st), "rev"))]
        tot = {k: 0.0 for k, _ in seq}
        reps = min(K, 5)
        launches_per_step, dirty_blocks = 0, None
        for rep in range(reps):
            p_holder["p"] = eng._params(eng.t, full=not eng.delta)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(seq) + 1)]
            c_before = lib.ppde_last_launch_count()
            evs[0].record()
            for i, (_, fn) in enumerate(seq):
                fn()
                evs[i + 1].record()
            torch.cuda.synchronize()
            launches_per_step = lib.ppde_last_launch_count() - c_before
            eng.t += 1
            for i, (k, _) in enumerate(seq):
                tot[k] += evs[i].elapsed_time(evs[i + 1])
            if eng.inc and rep == reps - 1:
                dm = eng.dmask.to(torch.int64) & 0xFFFFFFFF
                dirty_blocks = float(sum(((dm >> q) & 1).sum() for q in range(32)).item())
        breakdown = {k: v / reps for k, v in tot.items()}
        if eng.inc and eng.delta:
            # the exact backward that replaces the delta backward every bwd_refresh-th iteration (same inputs, same outputs)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(); eng.cnn_backward_y(st, do_fit=False, parts=7, full=True); ev1.record()
            torch.cuda.synchronize()
            refresh_ms = ev0.elapsed_time(ev1)
        P, Cc = L - 4, L
        per_sum = " (every kernel timed alone with CUDA events inside an eager replay of the step)"
        if eng.inc:
            # roofline of the DOMINANT kernel of the step (largest live duration); algorithmic work per SURVEY.md §8d
            NE = 20 * L
            tflop = lambda fl, ms_: fl / (ms_ * 1e-3) / 1e12
            gbs = lambda by, ms_: by / (ms_ * 1e-3) / 1e9
            # `traffic` = dram__bytes_read + dram__bytes_write of one launch from the committed `ncu --set full` captures of
            # round 2 (profiles/r02_ncu_full_summary_v49.txt, 16,384 chains in steady state, scaled to this launch)
            nd_ = (dirty_blocks / n) if dirty_blocks is not None else 2.25
            cand = {
                "cnn_backward_tc": dict(kernel=("cnn_backward_delta_kernel" if eng.delta else "cnn_backward_tc_kernel"), bound="tensor", unit="TFLOP/s",
                                        work=3 * (4 * Cc * Cc + 200 * P * Cc) * n, peak=pk["bf16_sustained"],
                                        traffic=((0.738e9 / 16384) if eng.delta else (660.6e6 / 8192)) * n if L == 238 else None,
                                        note="gradient of the CNN ensemble for every proposal: 3*(4C^2+200PC) algorithmic flops per chain; "
                                             "fp16 hi/lo split = 3 tensor-core passes per flop; " +
                                             ("delta mode: one tile of the touched positions per chain and net, only the adjoint rows that differ are gathered "
                                              "(~60 W1 rows of 960 B per chain and net from L2); bound by the producers' instruction streams, not by the tensor pipe" if eng.delta else
                                              "limited by the L2->SM gather of the winners' W1 rows (422 KB per chain and net)")),
                "cnn_forward_inc_tc": dict(kernel="cnn_forward_inc_kernel", bound="tensor", unit="TFLOP/s",
                                           # EXECUTED algorithmic flops: only the dirty 16-position blocks are evaluated (the full-evaluation
                                           # equivalent, 3*2*P*C*2C per chain, is reported under roofline.forward_incremental)
                                           work=nd_ * n * 3 * 2 * m.PB * Cc * 2 * Cc, peak=pk["bf16_sustained"],
                                           traffic=(0.505e9 / 16384) * n if L == 238 else None,
                                           note="max-pool winners of every proposal: 3 nets * 2*PB*C*2C flops per dirty PB-position block (PB = 8) "
                                                "(3 fp16 passes per flop); bound by the shared-memory pipe of the r1 producers (5 table reads per element)"),
                "cnn_inc_merge": dict(kernel="cnn_inc_merge_kernel", bound="hbm", unit="GB/s",
                                      # per channel: 16 B top-2 list read + 8 B per dirty block key + 16 B list + 8 B winner written
                                      work=int(n * 3 * 2 * Cc * (16 + 8 * nd_ + 24)), peak=pk["hbm_gbs"],
                                      traffic=(2.627e9 / 16384) * n if L == 238 else None,
                                      note="chain-level winner from the row's top-2 list and the dirty blocks' keys (exact, csrc/cnn_tc.cu)"),
                "pas_propose": dict(kernel="pas_propose_pos_kernel", bound="hbm", unit="GB/s",
                                    work=((4 * NE + L) + ((8 * m.D) if p_holder["p"].fuse_potts else 0)) * n, peak=pk["hbm_gbs"],
                                    traffic=(1.286e9 / 16384) * n if L == 238 else None,
                                    note="reads one gradient row per chain (and, fused, reads / writes the chain's Potts field rows); bound by instruction "
                                         "issue (Philox4x32-10 + race test for each of the 20L entries of every live sub-step: torch.multinomial needs one "
                                         "uniform per entry), not by bytes"),
            }
            dom = max(cand, key=lambda k_: breakdown[k_])
            cd = cand[dom]
            ach = tflop(cd["work"], breakdown[dom]) if cd["unit"] == "TFLOP/s" else gbs(cd["work"], breakdown[dom])
            roof = {"kernel": cd["kernel"], "bound": cd["bound"], "achieved": ach, "peak": cd["peak"], "unit": cd["unit"],
                    "frac": ach / cd["peak"], "traffic": cd["traffic"],
                    "peak_source": pk["src"] + (", sustained bf16 (kernel timed inside a long step)" if cd["bound"] == "tensor" else ", HBM copy"),
                    ("algorithmic_flops_per_launch" if cd["bound"] == "tensor" else "algorithmic_bytes_per_launch"): cd["work"],
                    "avg_launch_ms": breakdown[dom], "note": cd["note"] + per_sum}
            if eng.delta:
                delta_ms = breakdown["cnn_winner_sort"] + breakdown["cnn_backward_tc"] + breakdown["cnn_grad_combine"]
                R_ = m.bwd_refresh
                # the timed K iterations contain ceil(K / R) exact-refresh iterations (see above); the long-run share is K / R
                over = (refreshes / K - 1.0 / R_) * (refresh_ms - delta_ms)          # ms per step charged beyond the long-run share
                roof["backward"] = {"mode": "delta", "exact_refresh_every": R_, "exact_backward_ms": refresh_ms,
                                    "delta_backward_ms": delta_ms, "refresh_iterations_in_timed_region": refreshes,
                                    "refresh_ms_per_step_long_run": (refresh_ms - delta_ms) / R_,
                                    "value_at_long_run_refresh_share": n * world / ((ms / K - over) * 1e-3)}
            # the incremental forward computes only the dirty 16-position blocks: 3 nets * 2*16*C*2C flops per block
            if dirty_blocks is not None:
                f_inc = dirty_blocks * 3 * 2 * m.PB * Cc * 2 * Cc
                t_inc = breakdown["cnn_forward_inc_tc"] * 1e-3
                roof["forward_incremental"] = {
                    "kernel": "cnn_forward_inc_kernel", "dirty_blocks_per_chain": dirty_blocks / n, "blocks_per_chain": m.NB, "positions_per_block": m.PB,
                    "executed_algorithmic_tflops": f_inc / t_inc / 1e12, "avg_launch_ms": breakdown["cnn_forward_inc_tc"],
                    "full_evaluation_equivalent_tflops": 3 * 2 * P * Cc * 2 * Cc * n / ((breakdown["cnn_dirty"] + breakdown["cnn_inc_scan"]
                                                         + breakdown["cnn_forward_inc_tc"] + breakdown["cnn_inc_merge"]) * 1e-3) / 1e12}
                # per channel: 16 B top-2 list + the dirty blocks' keys read; 16 B list + 8 B winner written
                mbytes = int(n * 3 * 2 * Cc * (16 + 8 * (dirty_blocks / n) + 24))
                roof["merge_kernel"] = {"kernel": "cnn_inc_merge_kernel", "bound": "hbm", "algorithmic_bytes": mbytes,
                                        "achieved_gbs": mbytes / (breakdown["cnn_inc_merge"] * 1e-3) / 1e9, "peak_gbs": pk["hbm_gbs"],
                                        "frac": mbytes / (breakdown["cnn_inc_merge"] * 1e-3) / 1e9 / pk["hbm_gbs"]}
        else:
            flops_fwd = 3 * 2 * P * Cc * 2 * Cc * n                    # 3*2*P*C*2C per chain
            t_fwd = breakdown["cnn_forward"] * 1e-3
            achieved = flops_fwd / t_fwd / 1e12
            roof = {"kernel": ("cnn_forward_tc2_kernel" if os.environ.get("PPDE_TC_CTAS", "2") != "1" else "cnn_forward_tc_kernel") if m.cnn_forward_impl == "tc" else "cnn_forward_kernel", "bound": "tensor", "achieved": achieved, "peak": pk["bf16_sustained"],
                    "unit": "TFLOP/s", "frac": achieved / pk["bf16_sustained"],
                    # dram__bytes_read+write of this kernel from the committed `ncu --set full` capture
                    # (profiles/r01_cnn_forward_tc2_v3_summary.txt: 3.7 + 221.8 MB at 8192 chains), scaled to this launch
                    "traffic": (225.5e6 / 8192) * n if m.cnn_forward_impl == "tc" and L == 238 else None,
                    "peak_source": pk["src"] + ", sustained bf16 (kernel timed inside a long step)",
                    "algorithmic_flops_per_launch": flops_fwd, "avg_launch_ms": breakdown["cnn_forward"]}
        hbm_bytes = (8 * m.D + 2 * L + 16) * n                        # SURVEY.md §8d B_alg per chain-step
        t_hbm = (breakdown["pas_propose"] + breakdown["potts_incremental"] + breakdown["pas_reverse_accept"]) * 1e-3
        pp_ = p_holder["p"]
        fused = bool(pp_.comb_nets > 0)
        # what these kernels move by design: the gradient row read by the proposal and by the reverse proposal (4 NE each), the
        # Potts field rows (8 D) and - with the gradient combine fused into the reverse kernel - the current state's gradient
        # row, both field rows again and the proposal's row written once (8 NE + 8 D; the sparse per-net changes not counted)
        design_bytes = (8 * m.D + 2 * L + 16 + 8 * 20 * L + ((8 * 20 * L + 8 * m.D - 4 * 20 * L) if fused else 0)) * n
        roof["hbm_kernels"] = {"achieved_gbs": hbm_bytes / t_hbm / 1e9, "peak_gbs": pk["hbm_gbs"],
                               "frac": hbm_bytes / t_hbm / 1e9 / pk["hbm_gbs"],
                               "kernels": "pas_propose (+ Potts field update) + pas_reverse_accept" + (" (+ gradient combine)" if fused else "")
                                          if pp_.fuse_potts else "pas_propose + potts_incremental + pas_reverse_accept",
                               "algorithmic_bytes_per_step": hbm_bytes,
                               "bytes_moved_by_design_per_step": design_bytes,
                               "frac_of_design_bytes": design_bytes / t_hbm / 1e9 / pk["hbm_gbs"],
                               "note": "algorithmic bytes = SURVEY.md 8d (Potts field update + state); the same kernels also read the gradient rows"
                                       " and (fused) assemble / write the proposal's gradient row: bytes_moved_by_design counts those too"}
    if launches_per_step is None:
        launches_per_step = 12 if m.cnn_inc else 9

    # ---- end to end through the reference-facing API, host buffers in / out ----------------------
    e2e = None
    if not args.no_e2e:
        import argparse as ap_
        import gc
        win_hi = pr["win_lo"] + pr["J"].shape[0] - 1
        log_every = args.log_every if args.log_every > 0 else 10 ** 9
        del eng          # its pools return to torch's caching allocator and are reused by the runs below (warm allocator, as in
                         # any long-lived process); every copy and every kernel of a call stays inside its timed region
        gc.collect()

        def e2e_call(residue_io):
            """One PPDE_PAS.run from HOST buffers to HOST results; returns (seconds, bytes in, bytes out, phases, reports)."""
            local = not args.global_population
            sargs = ap_.Namespace(ppde_pas_length=wl["pas"], nmut_threshold=wl["nmut"], paper_results=wl["paper"], seed=0,
                                  ppde_verbose=False, ppde_local_population=local, ppde_residue_io=residue_io)
            smp = PPDE_PAS(sargs)
            n_in = n if local else n * world
            if residue_io:   # opt-in boundary format: uint8 residue indices [n, L], 1 byte per residue
                pop_host = torch.from_numpy(np.tile(pr["wt"], (n_in, 1))).pin_memory()
            else:            # the reference's format: float one-hot [n, L, 20] in host memory
                pop_host = torch.nn.functional.one_hot(torch.from_numpy(pr["wt"].astype(np.int64)), 20).float()[None] \
                    .repeat(n_in, 1, 1).pin_memory()
            gc.collect()
            gc.disable()     # no generational collection in the middle of the timed call
            barrier()
            t0 = time.perf_counter()
            out = smp.run(pop_host, K, energy, pr["win_lo"], win_hi, None, log_every=log_every)
            best = out[0]                                          # host tensor already (same device as the input)
            assert best.device.type == "cpu"
            barrier()
            dt = time.perf_counter() - t0
            gc.enable()
            if world > 1:
                tt = torch.tensor([dt], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            # bytes that actually crossed PCIe: residues in, residues + histories + best energies out
            h2d = n_in * L * world
            d2h = (n_in * m.aa_stride + out[3].nbytes + out[4].nbytes + out[1].nbytes + out[2].nbytes) * world
            rep_ms = [round(r_["device_ms"], 3) for _, r_ in smp.reports]
            return dt, h2d, d2h, dict(getattr(smp, "last_phases", {})), rep_ms

        # FIRST call of each variant is the reported one (no best-of); a second call is recorded beside it
        calls = {}
        variants = [("residue_io", True)]
        if not (args.global_population and world > 1):     # a global float one-hot is 80 bytes per residue on EVERY rank's host
            variants.append(("onehot_host", False))
        for variant, rio in variants:
            r1 = e2e_call(rio)
            r2 = e2e_call(rio)
            calls[variant] = (r1, r2)
        dt, h2d, d2h, phases, rep_ms = calls["residue_io"][0]
        dt_oh = calls["onehot_host"][0][0] if "onehot_host" in calls else None
        e2e = {"value": n * world * K / dt, "unit": UNIT, "h2d_bytes_per_step": h2d // K, "d2h_bytes_per_step": d2h // K,
               "what": "PPDE_PAS.run(host population -> 6-tuple on the host), K iterations incl. engine set-up and the t=0 "
                       "evaluation; wall clock of the FIRST call; population crosses the API as uint8 residue indices "
                       "(args.ppde_residue_io, INTEGRATION.md)",
               "calls_ms": [round(1e3 * c_[0], 1) for c_ in calls["residue_io"]],
               "population": "global on every rank (reference convention)" if args.global_population else "one shard per rank",
               # log_every population reports inside the call: device time of each (per-rank kernels + NCCL all-gathers /
               # all-reduce over NVLink + selection / unique-count kernels on the gathered vectors)
               "log_reports_in_call": len(rep_ms), "log_report_device_ms": rep_ms,
               "host_phases_ms": {k_: round(v_, 2) for k_, v_ in phases.items()}}
        if dt_oh is not None:
            # the reference's own boundary format: float one-hot [n, L, 20] in HOST memory in and out; the host cores reduce it to
            # residues before the copy and expand best_x after it (ppde_host_onehot_to_aa / ppde_host_aa_to_onehot)
            e2e["onehot_host_api"] = {"value": n * world * K / dt_oh, "calls_ms": [round(1e3 * c_[0], 1) for c_ in calls["onehot_host"]],
                                      "host_onehot_bytes_per_call": 2 * n * L * 20 * 4 * world,
                                      "host_phases_ms": {k_: round(v_, 2) for k_, v_ in calls["onehot_host"][0][3].items()}}

    # ---- CPU baseline (the unmodified reference on host cores when staged, else the port), rank 0, N=1 only -----------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = use_all_host_threads()
        v, _, kind, what = cpu_reference_throughput(wl, pr, args.cpu_chains, args.cpu_steps, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{args.cpu_chains} chains x {args.cpu_steps} iterations of the same workload "
                         f"(L={L}, pas={wl['pas']}); {what}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": args.workload, "desc": wl["desc"], "chains_per_gpu": n, "L": L,
                           "sub_steps": 2 * wl["pas"] - 1, "lamda": wl["lamda"], "nmut_threshold": wl["nmut"],
                           "l2": "per-step working set (gradient rows, >2 GB at 64k chains) exceeds the 126 MB L2; no flush needed",
                           "exact_refresh_iterations_in_timed_region": refreshes,
                           "parallelism": f"chains sharded x{world}, weights replicated, no collective in the step"},
                "clocks": clk, "e2e": e2e, "gpu_launches": launches_per_step * K,
                "roofline": roof, "cpu_baseline": cpu,
                "kernel_ms": breakdown, "accept_rate_last_step": acc_rate, "mean_edit_distance": mean_dist}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
