"""PPDE path-auxiliary sampler with the reference's interface (ppde/protein_samplers/ppde.py:8-192).

    sampler = PPDE_PAS(args)          # args.ppde_pas_length, args.nmut_threshold, args.paper_results [, args.seed]
    best_x, best_energy, best_fitness, energy_history, fitness_history, random_traj = \
        sampler.run(initial_population, num_steps, energy_function, min_pos, max_pos, oracle, log_every)

Differences from the reference that are deliberate and documented in DESIGN.md:
  * randomness comes from indexed Philox streams keyed by (seed, t, sub-step, global chain, entry)
    instead of the global torch/numpy generators (ppde_b200/philox.py);
  * the energy/gradient at the current state is cached from the previous iteration (Potts field
    updated incrementally) instead of being recomputed;
  * the population is not copied to the host every iteration (ppde.py:142,146): best-of-history and
    the random trajectory are tracked on the device with the same first-maximum rule (ppde.py:173);
  * under torch.distributed each rank runs its contiguous block of chains; the returned 6-tuple is
    the whole population on every rank.
"""
import os
import time

import numpy as np
import torch

from . import dist as D
from .energy import as_b200_energy
from .engine import ChainEngine


class _Phases:
    """Wall-clock phase times of one `run` call (PPDE_TRACE=1): a device synchronize closes every phase, so it is a
    diagnostic, not something to leave on when measuring throughput."""

    def __init__(self, device):
        self.mode = os.environ.get("PPDE_TRACE", "0")          # "1": synchronize at every mark and print; "2": print host-side times
        self.on = True                                         # host-side times are always recorded (PPDE_PAS.last_phases)
        self.device = device
        self.rows = []
        self.t = time.perf_counter()

    def mark(self, name):
        if self.on:
            if self.mode == "1":
                torch.cuda.synchronize(self.device)
            now = time.perf_counter()
            self.rows.append((name, (now - self.t) * 1e3))
            self.t = now

    def report(self):
        if self.mode in ("1", "2"):
            print("[ppde trace] " + "  ".join(f"{k}={v:.1f}ms" for k, v in self.rows), flush=True)


class PPDE_PAS:
    def __init__(self, args):
        self.ppde_temp = 2                                   # ppde.py:11 (g(t) = sqrt(t)); baked into the kernels
        self.ppde_pas_length = args.ppde_pas_length
        self.nmut_threshold = args.nmut_threshold
        self.paper_results = args.paper_results
        if self.nmut_threshold == 0:
            self.nmut_threshold = np.iinfo(np.int32).max      # ppde.py:15-17
        self.seed = int(getattr(args, "seed", 0))
        self.use_graph = bool(getattr(args, "ppde_cuda_graph", True))
        self.verbose = bool(getattr(args, "ppde_verbose", True))
        # True: `initial_population` is this rank's shard (equal sizes on every rank) and the returned
        # 6-tuple covers the local chains only; False (reference behaviour): global population in and out.
        self.local_population = bool(getattr(args, "ppde_local_population", False))
        # opt-in: the population crosses the API as residue indices uint8 [n, L] (alphabet order
        # ACDEFGHIKLMNPQRSTVWY, ppde/third_party/hsu/data_utils.py:48-72) instead of the float one-hot [n, L, 20], and best_x
        # comes back in the same form - the lossless 1-byte form of the state, 80x fewer bytes than the one-hot
        self.residue_io = bool(getattr(args, "ppde_residue_io", False))
        self.engine = None
        self.top_k = int(getattr(args, "ppde_top_k", 16))
        self.reports = []                                    # (iteration, report dict) of every log_every report

    def approximate_energy_change(self, score_change):       # ppde.py:20-21
        return score_change / self.ppde_temp

    def _print(self, *a, **k):
        if self.verbose and D.world()[0] == 0:
            print(*a, **k)

    def run(self, initial_population, num_steps, energy_function, min_pos, max_pos, oracle, log_every=50):
        """initial_population: float one-hot [n_chains, L, 20] (the GLOBAL population on every rank), on the device or in
        host memory (a host one-hot is reduced to residue indices by the host cores before the copy); with
        `args.ppde_residue_io` uint8 residue indices [n_chains, L].  best_x comes back in the same form on the same device."""
        self._print(min_pos, max_pos)
        energy = as_b200_energy(energy_function)
        m = energy.model
        rank, ws = D.world()
        if self.local_population:
            n = int(initial_population.size(0)) * ws
            lo, hi = D.shard_range(n, rank, ws)
            pop_local = initial_population
        else:
            n = int(initial_population.size(0))
            lo, hi = D.shard_range(n, rank, ws)
            pop_local = initial_population[lo:hi]
        random_idx = np.random.randint(0, n)                  # ppde.py:37 (numpy global generator)
        random_idx = D.agree_int(random_idx, 0, m.device)     # ranks must track the same chain
        own_rank, own_local = D.owner_of(random_idx, n, ws)
        thr = 0 if self.nmut_threshold == np.iinfo(np.int32).max else self.nmut_threshold

        ph = _Phases(m.device)
        with torch.cuda.device(m.device):
            aa0 = self._residues_on_device(m, pop_local)
            ph.mark("onehot_to_aa")
            eng = ChainEngine(m, hi - lo, self.ppde_pas_length, thr, self.paper_results, seed=self.seed,
                              chain_offset=lo, num_steps=num_steps,
                              traj_chain=own_local if own_rank == rank else -1,
                              min_pos=int(min_pos), max_pos=int(max_pos))
            eng.init_population(aa0)
            self.engine = eng
            ph.mark("engine_init")

            self.reporter = D.PopulationReporter(m.lib, m.device, n, m.L, m.aa_stride, m.wt,
                                                 top_k=int(getattr(self, "top_k", 16)))
            self._chain_lo = lo
            # iteration-0 report (ppde.py:48-57), quantiles by the device-side selection kernels
            gt = self._oracle_scores(oracle, eng, m) if oracle is not None else None
            rep = self.reporter.report(eng.E_hist[0], eng.fit_hist[0], gt, None, eng.aa, lo, want_topk=False)
            self.last_report = rep
            eq, fq = rep["energy_q"], rep["fitness_q"]
            self._print(f'[Iteration 0] energy: 50% {eq[0]:.3f}, 90% {eq[1]:.3f}')
            self._print(f'[Iteration 0] pred fit 50% {fq[0]:.3f}, 90% {fq[1]:.3f}')
            if oracle is not None:
                gq = rep["oracle_q"]
                self._print(f'[Iteration 0] oracle fit 50% {gq[0]:.3f}, 90% {gq[1]:.3f}')
            self._print('')
            ph.mark("report0")

            t = 0
            while t < num_steps:
                # next iteration i >= t that logs: i > 0 and (i + 1) % log_every == 0  (ppde.py:155)
                i = max(t, 1)
                i += (-(i + 1)) % log_every
                stop = min(i + 1, num_steps)
                eng.run_steps(stop - t, use_graph=self.use_graph)
                t = stop
                if t == i + 1:
                    self._log(eng, m, oracle, i)

            ph.mark("steps")
            # final 6-tuple (ppde.py:172-192)
            if self.local_population:
                gat = lambda t, dim=0: t
            else:
                gat = lambda t, dim=0: D.all_gather_cat(t, n, dim=dim)
            best_x = self._population_out(m, gat(eng.best_aa), initial_population.device)
            best_e, best_f = gat(eng.best_E).cpu().numpy(), gat(eng.best_fit).cpu().numpy()
            e_hist = gat(eng.E_hist, 1).cpu().numpy()
            f_hist = gat(eng.fit_hist, 1).cpu().numpy()
            traj = eng.traj_aa if eng.traj_aa is not None else torch.zeros(
                num_steps + 1, m.aa_stride, dtype=torch.uint8, device=m.device)
            traj = D.broadcast_from(traj, own_rank)
            random_traj = list(m.aa_to_onehot(traj).cpu().numpy())
            ph.mark("results")
            ph.report()
            self.last_phases = dict(ph.rows)
        return best_x, best_e, best_f, e_hist, f_hist, random_traj

    # -- the population across the API boundary ----------------------------------------------------------
    @staticmethod
    def _host_threads():
        try:
            cores = len(os.sched_getaffinity(0))
        except Exception:
            cores = os.cpu_count() or 1
        local_ws = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        return max(1, min(32, cores // max(local_ws, 1)))

    @staticmethod
    def _pinned(m, key, shape, dtype):
        """Pinned staging buffers are cached on the model (cudaHostAlloc costs milliseconds per call)."""
        cache = m.__dict__.setdefault("_pinned_cache", {})
        t = cache.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype).pin_memory()
            cache[key] = t
        return t

    def _residues_on_device(self, m, pop):
        """The caller's population as uint8 residues [n, aa_stride] on the model's device."""
        n = int(pop.shape[0])
        if self.residue_io:
            if pop.dtype != torch.uint8 or pop.dim() != 2 or pop.shape[1] != m.L:
                raise ValueError(f"ppde_residue_io: expected uint8 residues [n, {m.L}], got {pop.dtype} {tuple(pop.shape)}")
            aa = torch.zeros(n, m.aa_stride, dtype=torch.uint8, device=m.device)
            aa[:, :m.L].copy_(pop, non_blocking=True)
            return aa
        if pop.dim() != 3 or pop.shape[1] != m.L or pop.shape[2] != 20:
            raise ValueError(f"expected one-hot [n, {m.L}, 20], got {tuple(pop.shape)}")
        if pop.device.type == "cpu":
            x = pop.detach().to(torch.float32).contiguous()
            stage = self._pinned(m, "in", (n, m.aa_stride), torch.uint8)
            rc = m.lib.ppde_host_onehot_to_aa(x.data_ptr(), n, m.L, stage.data_ptr(), m.aa_stride, self._host_threads())
            if rc != 0:
                raise RuntimeError(f"ppde_host_onehot_to_aa failed with {rc}")
            return stage.to(m.device, non_blocking=True)
        return m.onehot_to_aa(pop)

    def _population_out(self, m, best_aa, device):
        """best_x in the form and on the device the population came in (ppde.py:191: `.to(initial_population.device)`)."""
        n = int(best_aa.shape[0])
        if self.residue_io:
            return best_aa[:, :m.L].to(device)
        if torch.device(device).type == "cpu":
            stage = self._pinned(m, "out", (n, m.aa_stride), torch.uint8)
            stage.copy_(best_aa)                                      # D2H of 1 byte per residue (synchronous: pinned target)
            x = torch.empty(n, m.L, 20, dtype=torch.float32)
            rc = m.lib.ppde_host_aa_to_onehot(stage.data_ptr(), m.aa_stride, n, m.L, x.data_ptr(), self._host_threads())
            if rc != 0:
                raise RuntimeError(f"ppde_host_aa_to_onehot failed with {rc}")
            return x
        return m.aa_to_onehot(best_aa).to(device)

    @staticmethod
    def _oracle_scores(oracle, eng, m):
        """oracle(cur_x) (ppde.py:48,156): our own model scores the residue states straight from the sampler's
        field rows; any other callable gets the reference's float one-hot view."""
        if hasattr(oracle, "score_engine") and getattr(oracle, "model", None) is m:
            return oracle.score_engine(eng).reshape(-1)
        return oracle(m.aa_to_onehot(eng.aa)).detach().float().reshape(-1)

    def _log(self, eng, m, oracle, i):
        """ppde.py:155-170: history row i+1 (pre-reset energies), post-reset states for the oracle and the distances."""
        gt = self._oracle_scores(oracle, eng, m) if oracle is not None else None
        rep = self.reporter.report(eng.E_hist[i + 1], eng.fit_hist[i + 1], gt, eng.accept, eng.aa, self._chain_lo)
        self.last_report = rep
        self.reports.append((i, rep))
        self._print(f'[Iteration {i}] energy: 50% {rep["energy_q"][0]:.3f}, 90% {rep["energy_q"][1]:.3f}', flush=True)
        self._print(f'[Iteration {i}] pred 50% {rep["fitness_q"][0]:.3f}, 90% {rep["fitness_q"][1]:.3f}', flush=True)
        if gt is not None:
            self._print(f'[Iteration {i}] oracle 50% {rep["oracle_q"][0]:.3f}, 90% {rep["oracle_q"][1]:.3f}', flush=True)
        self._print(f'   # accepted = {rep["accepted"]}')
        self._print(f'   # dist = {rep["mean_dist"]}')
        self._print(f'   # diversity = {rep["diversity_pct"]:.1f}%')
        self._print('', flush=True)
