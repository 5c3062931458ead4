"""Chain sharding across ranks and the log_every / end-of-run population gathers.

Chains are independent (no chain reads another chain's state anywhere in
ppde/protein_samplers/ppde.py:65-153), so each rank owns a contiguous block of chains and the
step loop contains NO collective.  Collectives (NCCL over NVLink on GPUs; gloo in the CPU
tests) appear only where the reference reduces over the whole population: the log_every
report (ppde.py:155-168) and the final 6-tuple (ppde.py:172-192).
Random streams are indexed by GLOBAL chain id, so results do not depend on the world size.
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank, world_size):
    """Contiguous block [lo, hi) of `n` chains owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n, world_size):
    return [shard_range(n, r, world_size)[1] - shard_range(n, r, world_size)[0] for r in range(world_size)]


def all_gather_cat(t, n_global, dim=0):
    """Concatenate per-rank shards (possibly of unequal size along `dim`) on every rank."""
    rank, ws = world()
    if ws == 1:
        return t
    sizes = shard_sizes(n_global, ws)
    mx = max(sizes)
    t = t.movedim(dim, 0).contiguous()
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = torch.empty((ws * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad)
    parts = [out[r * mx:r * mx + sizes[r]] for r in range(ws)]
    return torch.cat(parts, 0).movedim(0, dim)


def all_reduce_sum(t):
    _, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_from(t, src):
    _, ws = world()
    if ws > 1:
        dist.broadcast(t, src=src)
    return t


def agree_int(value, src=0, device=None):
    """The same integer on every rank (rank `src`'s).  Host-side draws such as the reference's
    `np.random.randint` for the tracked chain (ppde.py:37) differ between processes unless every
    rank seeded numpy identically; collectives that depend on them (broadcast source) must agree."""
    _, ws = world()
    if ws == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=device if device is not None else "cpu")
    dist.broadcast(t, src=src)
    return int(t.item())


def owner_of(chain, n, world_size):
    for r in range(world_size):
        lo, hi = shard_range(n, r, world_size)
        if lo <= chain < hi:
            return r, chain - lo
    raise ValueError("chain outside the population")


class PopulationReporter:
    """The reference's log_every report (ppde.py:155-170) plus the population metrics of scripts/make_figures.py:29-49 and a
    torch.topk of the energies, computed ON THE DEVICE: per-rank kernels, NCCL all-gathers / one all-reduce of the shards
    (the only collectives of a run), exact selection / counting kernels on the gathered vectors, and ONE device->host copy
    of the finished numbers.  Nothing here touches the host except that last copy.

    Collectives per report: all-gather of [3, n/ws] floats (energy, predicted fitness, oracle fitness), all-gather of the
    residue states u8 [n/ws, stride] (diversity), all-reduce of 4 int64 (accepted, sum dist, sum dist^2, n), all-gathers
    of the k local top-k candidates (values, global chain ids, sequences).
    """
    Q = (0.5, 0.9)

    def __init__(self, lib, device, n_global, L, aa_stride, wt_dev, top_k=16):
        import ctypes as C
        self.C = C
        self.lib, self.dev = lib, device
        self.n, self.L, self.stride, self.wt = int(n_global), int(L), int(aa_stride), wt_dev
        _, ws = world()
        sizes = shard_sizes(self.n, ws)
        self.k = max(0, min(int(top_k), min(sizes), 1024))
        self.q = torch.tensor(self.Q, dtype=torch.float64, device=device)
        entries = int(lib.ppde_unique_count_table_entries(self.n))
        self.table = torch.empty(entries, dtype=torch.int32, device=device)
        self.out = torch.zeros(16, dtype=torch.float64, device=device)       # quantiles 0..5 | sums 6..9 | unique 10
        self.sums = torch.zeros(4, dtype=torch.int64, device=device)
        self.count = torch.zeros(1, dtype=torch.int32, device=device)

    def _p(self, t):
        return self.C.c_void_p(t.data_ptr()) if t is not None else self.C.c_void_p(0)

    def _st(self):
        return self.C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _check(self, code, what):
        if code != 0:
            raise RuntimeError(f"{what} failed with cudaError_t {code}")

    def report(self, energy, fitness, oracle_fit, accept, aa, chain_lo, want_topk=True):
        """energy / fitness / oracle_fit: float32 [n_local] (oracle_fit may be None); accept: uint8 [n_local] or None;
        aa: uint8 [n_local, stride] current states; chain_lo: global id of local chain 0.  Returns a dict of host numbers
        (and, with want_topk, 'topk_energy' / 'topk_chain' / 'topk_aa' as host arrays)."""
        lib, dev, st = self.lib, self.dev, self._st()
        nl = int(energy.shape[0])
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        # -- per-rank kernels ----------------------------------------------------------------------
        dist_l = torch.empty(nl, dtype=torch.int32, device=dev)
        self._check(lib.ppde_population_metrics(self._p(aa), self.stride, nl, self.L, self._p(self.wt), self._p(dist_l),
                                                self._p(None), st), "population_metrics")
        self._check(lib.ppde_population_sums(self._p(accept), self._p(dist_l), nl, self._p(self.sums), st), "population_sums")
        packed = torch.stack([energy.float(), fitness.float(),
                              oracle_fit.float() if oracle_fit is not None else torch.zeros_like(energy, dtype=torch.float32)])
        # -- the collectives -----------------------------------------------------------------------
        allv = all_gather_cat(packed, self.n, dim=1).contiguous()             # [3, n]
        all_aa = all_gather_cat(aa, self.n)                                   # [n, stride]
        all_reduce_sum(self.sums)
        # -- exact order statistics / counts on the gathered vectors ---------------------------------
        for j in range(3 if oracle_fit is not None else 2):
            self._check(lib.ppde_quantiles(self._p(allv[j]), self.n, self._p(self.q), 2,
                                           self.C.c_void_p(self.out.data_ptr() + 16 * j), st), "quantiles")
        self._check(lib.ppde_unique_count(self._p(all_aa), self.stride, self.n, self.L, self._p(self.table),
                                          self.table.numel(), self._p(self.count), st), "unique_count")
        self.out[6:10] = self.sums.to(torch.float64)
        self.out[10] = self.count.to(torch.float64)[0]
        top = None
        if want_topk and self.k > 0:
            k = self.k
            vals = torch.empty(k, dtype=torch.float32, device=dev)
            ids = torch.empty(k, dtype=torch.int64, device=dev)
            seqs = torch.empty(k, self.stride, dtype=torch.uint8, device=dev)
            self._check(lib.ppde_topk(self._p(energy), nl, k, int(chain_lo), self._p(None), self._p(vals), self._p(ids), st), "topk")
            self._check(lib.ppde_gather_rows(self._p(aa), self.stride, self._p(ids), k, int(chain_lo), self._p(seqs), st), "gather_rows")
            _, ws = world()
            if ws > 1:
                cv, ci, cs = gather_equal(vals), gather_equal(ids), gather_equal(seqs)     # [ws*k] candidates, rank-major
                pos = torch.empty(k, dtype=torch.int64, device=dev)
                fv = torch.empty(k, dtype=torch.float32, device=dev)
                self._check(lib.ppde_topk(self._p(cv), ws * k, k, 0, self._p(None), self._p(fv), self._p(pos), st), "topk(final)")
                vals, ids, seqs = fv, ci[pos], cs[pos]
            top = (vals, ids, seqs)
        ev1.record()
        host = self.out.cpu().numpy()                                           # the one D2H of the report
        n = float(host[9])
        mean_d = host[7] / n
        rep = {
            "energy_q": host[0:2].copy(), "fitness_q": host[2:4].copy(),
            "oracle_q": host[4:6].copy() if oracle_fit is not None else None,
            "accepted": float(host[6]), "mean_dist": float(mean_d),
            "std_dist": float(np.sqrt(max(host[8] / n - mean_d * mean_d, 0.0))),      # n_hops: population std (np.std)
            "unique": int(host[10]), "diversity_pct": float(host[10]) / n * 100.0,
            "device_ms": float(ev0.elapsed_time(ev1)),        # kernels + collectives of this report, on the device
        }
        if top is not None:
            rep["topk_energy"] = top[0].cpu().numpy()
            rep["topk_chain"] = top[1].cpu().numpy()
            rep["topk_aa"] = top[2][:, :self.L].cpu().numpy()
        return rep


def gather_equal(t):
    """all-gather of equally sized shards, concatenated along dim 0 (rank-major)."""
    _, ws = world()
    if ws == 1:
        return t
    out = torch.empty((ws * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous())
    return out
