"""Incremental CNN forward (dirty 16-position blocks only, block keys cached in a row pool) against the full
tensor-core forward: max-pool winners `mkey` and the relu-mask rows must be IDENTICAL bit for bit, for any set of
changed residues (none, one at either end, a few, more than 8 dirty blocks, everything)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ppde_port as port

pytestmark = pytest.mark.gpu


def _mutants(rng, x, L):
    """One proposal per chain: a mix of the patterns the sampler produces and the edge cases."""
    n = x.shape[0]
    y = x.copy()
    for b in range(n):
        kind = b % 8
        if kind == 0:
            continue                                     # no-op proposal: no dirty block
        if kind == 1:
            pos = np.array([0])
        elif kind == 2:
            pos = np.array([L - 1])
        elif kind == 3:
            pos = rng.integers(0, L, size=1)
        elif kind == 4:
            pos = rng.integers(0, L, size=3)
        elif kind == 5:
            pos = rng.integers(0, L, size=19)            # pas=10 paths: usually more than 8 dirty blocks
        elif kind == 6:
            pos = np.arange(L)                           # everything changes
        else:
            pos = np.array([15, 16, 31])                 # block boundaries
            pos = pos[pos < L]
        y[b, pos] = (x[b, pos] + rng.integers(1, 20, size=pos.shape[0])) % 20
    return y


@pytest.mark.parametrize("L,n", [(40, 24), (104, 40), (237, 48), (238, 331)])
def test_incremental_forward_is_bit_identical(L, n):
    from ppde_b200 import _lib
    from ppde_b200.engine import PoEModel, _ptr, _stream
    w = port.synthetic_weights(L, seed=L + 7, lamda=1.0)
    m = PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    if not m.cnn_inc:
        pytest.skip("incremental path disabled")
    lib, dev = m.lib, m.device
    rng = np.random.default_rng(L)
    x = np.tile(w.wt, (n, 1)).astype(np.uint8)
    for b in range(n):
        pos = rng.integers(0, L, size=b % 11)
        x[b, pos] = rng.integers(0, 20, size=pos.shape[0])
    y = _mutants(rng, x, L)

    def dev_aa(a):
        pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = a
        return torch.from_numpy(pad).to(dev)
    ax, ay = dev_aa(x), dev_aa(y)
    J2, nets, P, NB = 2 * m.C, m.n_nets, m.P, m.NB
    rows = 2 * n + 1
    bkey = torch.full((rows * nets * NB * J2,), -1, dtype=torch.int64, device=dev)
    r1pool = torch.full((rows * nets * P * 32,), 0xAB, dtype=torch.uint8, device=dev)
    st = _stream()
    # current states: private row b for even chains, row n + b for odd ones (both halves of the pool get used)
    rows_x = torch.tensor([b if b % 2 == 0 else n + b for b in range(n)], dtype=torch.int32, device=dev)
    rows_y = torch.tensor([n + b if b % 2 == 0 else b for b in range(n)], dtype=torch.int32, device=dev)
    mk_x = torch.zeros(n * nets * J2, dtype=torch.int64, device=dev)
    m.cnn_forward_pool(ax, n, mk_x, bkey, r1pool, None, None, rows_x, 0, st)         # full evaluation into rows_x
    # full tensor-core kernel on x and y (the reference for bit-exactness)
    mk_fx = torch.zeros_like(mk_x); mk_fy = torch.zeros_like(mk_x)
    rm_fx = torch.zeros(n * nets * P * 32, dtype=torch.uint8, device=dev); rm_fy = torch.zeros_like(rm_fx)
    _lib.check(lib.ppde_cnn_forward_tc(C.byref(m.cnn), _ptr(ax), m.aa_stride, n, _ptr(mk_fx), _ptr(rm_fx), st), "tc x")
    _lib.check(lib.ppde_cnn_forward_tc(C.byref(m.cnn), _ptr(ay), m.aa_stride, n, _ptr(mk_fy), _ptr(rm_fy), st), "tc y")
    torch.cuda.synchronize()
    assert torch.equal(mk_x, mk_fx), "full evaluation through the block-key kernel differs from the full kernel"
    pool = r1pool.view(rows, nets * P * 32)
    assert torch.equal(pool[rows_x.long()], rm_fx.view(n, -1)), "relu-mask rows of the full evaluation differ"

    # incremental: dirty blocks of y against x
    dmask = torch.zeros(n, dtype=torch.int32, device=dev)
    _lib.check(lib.ppde_cnn_dirty(C.byref(m.cnn), _ptr(ax), _ptr(ay), m.aa_stride, n, _ptr(dmask), _ptr(r1pool),
                                  _ptr(rows_x), _ptr(rows_y), st), "dirty")
    mk_y = torch.zeros_like(mk_x)
    m.cnn_forward_pool(ay, n, mk_y, bkey, r1pool, dmask, rows_x, rows_y, 0, st)
    torch.cuda.synchronize()
    # dirty masks against a host restatement
    dm = dmask.cpu().numpy().astype(np.uint32)
    for b in range(n):
        want = 0
        for i in np.nonzero(x[b] != y[b])[0]:
            for p in range(max(i - 4, 0), min(i, P - 1) + 1):
                want |= 1 << (p >> 4)
        assert dm[b] == want, f"chain {b}: dirty mask {dm[b]:#x} != {want:#x}"
    assert (dm[::8] == 0).all() and (dm[6::8] == (1 << NB) - 1).all()
    assert torch.equal(mk_y, mk_fy), "incremental winners differ from the full kernel"
    assert torch.equal(pool[rows_y.long()], rm_fy.view(n, -1)), "incremental relu-mask rows differ"
    # the proposal rows now hold a complete cache: a second incremental step from y back to x reproduces x
    _lib.check(lib.ppde_cnn_dirty(C.byref(m.cnn), _ptr(ay), _ptr(ax), m.aa_stride, n, _ptr(dmask), _ptr(r1pool),
                                  _ptr(rows_y), _ptr(rows_x), st), "dirty back")
    mk_b = torch.zeros_like(mk_x)
    m.cnn_forward_pool(ax, n, mk_b, bkey, r1pool, dmask, rows_y, rows_x, 0, st)
    torch.cuda.synchronize()
    assert torch.equal(mk_b, mk_fx), "second incremental step (y -> x) differs from the full kernel"
    assert torch.equal(pool[rows_x.long()], rm_fx.view(n, -1))
