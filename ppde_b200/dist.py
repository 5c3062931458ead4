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


def population_report(energy, fitness, oracle_fit, accepted, edit_dist, seq_hash=None):
    """The reference's log_every report (ppde.py:158-168) from whole-population vectors (host numpy),
    plus diversity (% unique sequences, scripts/make_figures.py:38-49) when hashes are given."""
    rep = {
        "energy_q": np.quantile(energy, [0.5, 0.9]),
        "fitness_q": np.quantile(fitness, [0.5, 0.9]),
        "oracle_q": np.quantile(oracle_fit, [0.5, 0.9]) if oracle_fit is not None else None,
        "accepted": float(np.sum(accepted)),
        "mean_dist": float(np.mean(edit_dist)),
    }
    if seq_hash is not None:
        rep["diversity_pct"] = len(np.unique(seq_hash)) / len(seq_hash) * 100.0
    return rep
