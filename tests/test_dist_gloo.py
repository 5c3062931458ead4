"""The N>1 host path (sharding + population gathers) on CPU with the gloo backend, world_size 2."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, ws, port, n, T, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from ppde_b200 import dist as D
    lo, hi = D.shard_range(n, rank, ws)
    glob_e = torch.arange(n, dtype=torch.float32) * 0.5
    glob_hist = torch.arange((T + 1) * n, dtype=torch.float32).reshape(T + 1, n)
    glob_aa = (torch.arange(n * 7) % 20).to(torch.uint8).reshape(n, 7)
    e = D.all_gather_cat(glob_e[lo:hi].clone(), n)
    hist = D.all_gather_cat(glob_hist[:, lo:hi].clone(), n, dim=1)
    aa = D.all_gather_cat(glob_aa[lo:hi].clone(), n)
    acc = D.all_reduce_sum(torch.tensor([float(hi - lo)]))
    owner, local = D.owner_of(n - 1, n, ws)
    traj = torch.full((T + 1, 4), float(rank)) if rank == owner else torch.zeros(T + 1, 4)
    traj = D.broadcast_from(traj, owner)
    ok = (torch.equal(e, glob_e) and torch.equal(hist, glob_hist) and torch.equal(aa, glob_aa)
          and float(acc) == n and torch.equal(traj, torch.full((T + 1, 4), float(owner))))
    # host-side draws differ between processes; the tracked-chain index must be made to agree (ppde.py:37)
    ok = ok and D.agree_int(100 + 17 * rank, 0) == 100 and D.agree_int(5 + rank, 1) == 6
    # equal-size gather used by the top-k candidates of the device-side report
    cand = D.gather_equal(torch.full((3,), float(rank)))
    ok = ok and torch.equal(cand, torch.tensor([0.0] * 3 + [1.0] * 3))
    ok = ok and D.shard_sizes(n, ws) == [hi_ - lo_ for lo_, hi_ in (D.shard_range(n, r, ws) for r in range(ws))]
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [9, 16])
def test_gather_paths_world_size_2(tmp_path, n):
    port = 29500 + (os.getpid() % 2000) + n
    mp.spawn(_worker, args=(2, port, n, 3, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"
