"""Incremental CNN forward (dirty 16-position blocks only, block keys cached in a row pool) against the full
tensor-core forward: max-pool winners `mkey` and the relu-mask rows must be IDENTICAL bit for bit, for any set of
changed residues (none, one at either end, a few, more than 8 dirty blocks, everything)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ppde_port as port

pytestmark = pytest.mark.gpu


def _mask_rows(r1pool, btab, rows, nets, P, NB, PB=None):
    """Relu-mask rows [n, nets, P, 32] of the pool rows `rows`, read through the block table."""
    PB = PB or (P + NB - 1) // NB                    # positions per block (8: NB = ceil(P / PB))
    n_rows = btab.numel() // NB
    pool = r1pool.view(n_rows, nets, P, 32)
    tab = btab.view(n_rows, NB).long()[rows.long()]                    # [n, NB]
    src = tab[:, (torch.arange(P, device=tab.device) // PB)]           # [n, P] source row of every position
    pidx = torch.arange(P, device=tab.device)[None, :].expand_as(src)
    return pool[src, :, pidx, :].permute(0, 2, 1, 3).contiguous()      # [n, P, nets, 32] -> [n, nets, P, 32]


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


@pytest.mark.parametrize("use_pool", [True, False])
@pytest.mark.parametrize("L,n", [(40, 24), (104, 40), (237, 48), (238, 331)])
def test_incremental_forward_is_bit_identical(L, n, use_pool):
    """use_pool: the pool of raw row winners is kept, so the merge reads the current row's winner + the dirty blocks' keys
    (the production path); without it the merge scans all NB block keys per channel.  Same bits either way."""
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
    btab = torch.full((rows * NB,), -1, dtype=torch.int32, device=dev)
    mkpool = torch.full((rows * nets * J2 * 2,), -1, dtype=torch.int64, device=dev) if use_pool else None
    st = _stream()
    # current states: private row b for even chains, row n + b for odd ones (both halves of the pool get used)
    rows_x = torch.tensor([b if b % 2 == 0 else n + b for b in range(n)], dtype=torch.int32, device=dev)
    rows_y = torch.tensor([n + b if b % 2 == 0 else b for b in range(n)], dtype=torch.int32, device=dev)
    mk_x = torch.zeros(n * nets * J2, dtype=torch.int64, device=dev)
    m.cnn_forward_pool(ax, n, mk_x, bkey, r1pool, None, None, rows_x, 0, st, btab=btab, mkpool=mkpool)         # full evaluation into rows_x
    # full tensor-core kernel on x and y (the reference for bit-exactness)
    mk_fx = torch.zeros_like(mk_x); mk_fy = torch.zeros_like(mk_x)
    rm_fx = torch.zeros(n * nets * P * 32, dtype=torch.uint8, device=dev); rm_fy = torch.zeros_like(rm_fx)
    _lib.check(lib.ppde_cnn_forward_tc(C.byref(m.cnn), _ptr(ax), m.aa_stride, n, _ptr(mk_fx), _ptr(rm_fx), None, st), "tc x")
    _lib.check(lib.ppde_cnn_forward_tc(C.byref(m.cnn), _ptr(ay), m.aa_stride, n, _ptr(mk_fy), _ptr(rm_fy), None, st), "tc y")
    torch.cuda.synchronize()
    assert torch.equal(mk_x, mk_fx), "full evaluation through the block-key kernel differs from the full kernel"
    tab = btab.view(rows, NB)
    assert torch.equal(tab[rows_x.long()], rows_x[:, None].expand(n, NB)), "a fully evaluated row must point at itself"
    assert torch.equal(_mask_rows(r1pool, btab, rows_x, nets, P, NB), rm_fx.view(n, nets, P, 32)), "relu-mask rows of the full evaluation differ"

    # incremental: dirty blocks of y against x
    dmask = torch.zeros(n, dtype=torch.int32, device=dev)
    _lib.check(lib.ppde_cnn_dirty(C.byref(m.cnn), _ptr(ax), _ptr(ay), m.aa_stride, n, _ptr(dmask), st), "dirty")
    mk_y = torch.zeros_like(mk_x)
    m.cnn_forward_pool(ay, n, mk_y, bkey, r1pool, dmask, rows_x, rows_y, 0, st, btab=btab, mkpool=mkpool)
    torch.cuda.synchronize()
    # dirty masks against a host restatement
    dm = dmask.cpu().numpy().astype(np.uint32)
    for b in range(n):
        want = 0
        for i in np.nonzero(x[b] != y[b])[0]:
            for p in range(max(i - 4, 0), min(i, P - 1) + 1):
                want |= 1 << (p // m.PB)
        assert dm[b] == want, f"chain {b}: dirty mask {dm[b]:#x} != {want:#x}"
    assert (dm[::8] == 0).all() and (dm[6::8] == (1 << NB) - 1).all()
    assert torch.equal(mk_y, mk_fy), "incremental winners differ from the full kernel"
    assert torch.equal(_mask_rows(r1pool, btab, rows_y, nets, P, NB), rm_fy.view(n, nets, P, 32)), "incremental relu-mask rows differ"
    # table of the proposal rows: own slot for the dirty blocks, the current row's entry for the clean ones
    dmt = torch.from_numpy(dm.astype(np.int64)).to(dev)
    dirty = ((dmt[:, None] >> torch.arange(NB, device=dev)[None, :]) & 1).bool()
    want_tab = torch.where(dirty, rows_y[:, None].expand(n, NB), tab[rows_x.long()])
    assert torch.equal(tab[rows_y.long()], want_tab)
    # the current rows are untouched: x can still be read back exactly
    assert torch.equal(_mask_rows(r1pool, btab, rows_x, nets, P, NB), rm_fx.view(n, nets, P, 32))
    # the proposal rows now hold a complete cache: a second incremental step from y back to x reproduces x
    _lib.check(lib.ppde_cnn_dirty(C.byref(m.cnn), _ptr(ay), _ptr(ax), m.aa_stride, n, _ptr(dmask), st), "dirty back")
    mk_b = torch.zeros_like(mk_x)
    m.cnn_forward_pool(ax, n, mk_b, bkey, r1pool, dmask, rows_y, rows_x, 0, st, btab=btab, mkpool=mkpool)
    torch.cuda.synchronize()
    assert torch.equal(mk_b, mk_fx), "second incremental step (y -> x) differs from the full kernel"
    assert torch.equal(_mask_rows(r1pool, btab, rows_x, nets, P, NB), rm_fx.view(n, nets, P, 32))
    assert torch.equal(_mask_rows(r1pool, btab, rows_y, nets, P, NB), rm_fy.view(n, nets, P, 32)), "y (now the current rows) was damaged"
    # a third step x -> y2 exercises slots that the other row still points at
    y2 = _mutants(np.random.default_rng(L + 99), x, L)
    ay2 = dev_aa(y2)
    mk_fy2 = torch.zeros_like(mk_x); rm_fy2 = torch.zeros_like(rm_fx)
    _lib.check(lib.ppde_cnn_forward_tc(C.byref(m.cnn), _ptr(ay2), m.aa_stride, n, _ptr(mk_fy2), _ptr(rm_fy2), None, st), "tc y2")
    _lib.check(lib.ppde_cnn_dirty(C.byref(m.cnn), _ptr(ax), _ptr(ay2), m.aa_stride, n, _ptr(dmask), st), "dirty 3")
    mk_y2 = torch.zeros_like(mk_x)
    m.cnn_forward_pool(ay2, n, mk_y2, bkey, r1pool, dmask, rows_x, rows_y, 0, st, btab=btab, mkpool=mkpool)
    torch.cuda.synchronize()
    assert torch.equal(mk_y2, mk_fy2), "third incremental step differs from the full kernel"
    assert torch.equal(_mask_rows(r1pool, btab, rows_y, nets, P, NB), rm_fy2.view(n, nets, P, 32))
    assert torch.equal(_mask_rows(r1pool, btab, rows_x, nets, P, NB), rm_fx.view(n, nets, P, 32)), "the current rows were damaged"


@pytest.mark.parametrize("L,n", [(40, 24), (104, 40), (238, 96)])
def test_delta_backward_matches_full_backward(L, n):
    """G(y) = G(x) + change of the few adjoint rows that differ (delta backward) against the full backward of y."""
    from ppde_b200 import _lib
    from ppde_b200.engine import PoEModel, _ptr, _stream
    w = port.synthetic_weights(L, seed=L + 3, lamda=4.0, window=(1, L - 2))
    m = PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    if not m.cnn_bwd_delta:
        pytest.skip("delta backward disabled")
    lib, dev = m.lib, m.device
    rng = np.random.default_rng(L + 1)
    x = np.tile(w.wt, (n, 1)).astype(np.uint8)
    for b in range(n):
        pos = rng.integers(0, L, size=b % 13)
        x[b, pos] = rng.integers(0, 20, size=pos.shape[0])
    y = _mutants(rng, x, L)
    y[n - 1, 5:11] = y[n - 1, 20:26]                     # repeated 5-mers: exact max-pool ties

    def dev_aa(a):
        pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = a
        return torch.from_numpy(pad).to(dev)
    ax, ay = dev_aa(x), dev_aa(y)
    J2, nets, P, NB = 2 * m.C, m.n_nets, m.P, m.NB
    rows = 2 * n + 1
    f32 = torch.float32
    G = torch.full((rows, m.NE), float("nan"), dtype=f32, device=dev)
    Gp = torch.zeros(rows, m.D, dtype=f32, device=dev)
    bkey = torch.zeros(rows * nets * NB * J2, dtype=torch.int64, device=dev)
    r1pool = torch.zeros(rows * nets * P * 32, dtype=torch.uint8, device=dev)
    mkpool = torch.zeros(rows * nets * J2 * 2, dtype=torch.int64, device=dev)
    btab = torch.full((rows * NB,), -1, dtype=torch.int32, device=dev)
    E = torch.zeros(n, dtype=f32, device=dev); fit = torch.zeros_like(E); Ep = torch.zeros_like(E)
    st = _stream()
    m.evaluate_into(ax, n, G, 0, Gp, 0, E, fit, Ep, bkey=bkey, r1pool=r1pool, mkpool=mkpool, btab=btab)      # rows 0..n-1 = states x
    rows_x = torch.arange(n, dtype=torch.int32, device=dev)
    rows_y = rows_x + n
    Ep_y = torch.zeros_like(E); E_y = torch.zeros_like(E); fit_y = torch.zeros_like(E)
    m.potts_full(ay, n, C.c_void_p(Gp.data_ptr() + n * m.D * 4), _ptr(Ep_y), st)
    dmask = torch.zeros(n, dtype=torch.int32, device=dev)
    _lib.check(lib.ppde_cnn_dirty(C.byref(m.cnn), _ptr(ax), _ptr(ay), m.aa_stride, n, _ptr(dmask), st), "dirty")
    mk = m.ws.mkey(n)
    m.cnn_forward_pool(ay, n, mk, bkey, r1pool, dmask, rows_x, rows_y, 0, st, mkpool=mkpool, btab=btab)
    m.cnn_backward_delta(ax, ay, n, mk, mkpool, _ptr(Gp), _ptr(Ep_y), _ptr(G), rows_x, rows_y, E_y, fit_y, r1pool, st, btab=btab)
    torch.cuda.synchronize()
    g_delta = G[n:2 * n].cpu().numpy()
    e_delta, f_delta = E_y.cpu().numpy(), fit_y.cpu().numpy()
    E2, fit2, G2, _ = m.energy(ay)                        # full forward + full backward of y
    torch.cuda.synchronize()
    g_full = G2.reshape(n, -1).cpu().numpy()
    assert np.isfinite(g_delta).all()
    scale = np.abs(g_full).max(axis=1, keepdims=True)
    err = np.max(np.abs(g_delta - g_full) / scale)
    assert err < 2e-6, f"delta vs full gradient, relative to each chain's max |g|: {err:.3e}"
    assert np.array_equal(f_delta, fit2.cpu().numpy()) and np.array_equal(e_delta, E2.cpu().numpy())
    # a second delta step back to x (rows swap roles) stays within rounding of the exact gradient of x
    Gx_exact = G[:n].clone()
    _lib.check(lib.ppde_cnn_dirty(C.byref(m.cnn), _ptr(ay), _ptr(ax), m.aa_stride, n, _ptr(dmask), st), "dirty back")
    m.cnn_forward_pool(ax, n, mk, bkey, r1pool, dmask, rows_y, rows_x, 0, st, mkpool=mkpool, btab=btab)
    m.cnn_backward_delta(ay, ax, n, mk, mkpool, _ptr(Gp), _ptr(Ep), _ptr(G), rows_y, rows_x, E, fit, r1pool, st, btab=btab)
    torch.cuda.synchronize()
    gx = G[:n].cpu().numpy(); gx0 = Gx_exact.cpu().numpy()
    err = np.max(np.abs(gx - gx0) / np.abs(gx0).max(axis=1, keepdims=True))
    assert err < 4e-6, f"x -> y -> x round trip of the delta backward: {err:.3e}"
