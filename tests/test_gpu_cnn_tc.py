"""tcgen05 CNN forward vs the fp32 SIMT forward (same C-ABI contract) and vs the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ppde_port as port

pytestmark = pytest.mark.gpu


def _decode(mk, n, nets, J2):
    k = mk[: n * nets * J2].cpu().numpy().astype(np.uint64).reshape(n, nets, J2)
    val = (k >> np.uint64(32)).astype(np.uint32).view(np.float32)
    pos = (np.uint64(0xFFFFFFFF) - (k & np.uint64(0xFFFFFFFF))).astype(np.int64)
    return val, pos


@pytest.mark.parametrize("L,n", [(40, 7), (96, 33), (104, 20), (238, 19)])
def test_tc_forward_matches_simt_and_oracle(L, n):
    from ppde_b200 import _lib
    from ppde_b200.engine import PoEModel, _ptr, _stream
    w = port.synthetic_weights(L, seed=L, lamda=1.0)
    m = PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    rng = np.random.default_rng(L)
    aa = np.tile(w.wt, (n, 1))
    for b in range(1, n):
        pos = rng.integers(0, L, size=min(b, L))
        aa[b, pos] = rng.integers(0, 20, size=pos.shape[0])
    aa[n - 1] = rng.integers(0, 20, size=L)
    pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = aa
    aad = torch.from_numpy(pad).to(m.device)
    J2, nets = 2 * m.C, m.n_nets
    mk1 = torch.zeros(n * nets * J2, dtype=torch.int64, device=m.device)
    mk2 = torch.full((n * nets * J2,), -1, dtype=torch.int64, device=m.device)
    lib = m.lib
    _lib.check(lib.ppde_cnn_forward(C.byref(m.cnn), _ptr(aad), m.aa_stride, n, _ptr(mk1), _stream()), "simt")
    _lib.check(lib.ppde_cnn_forward_tc(C.byref(m.cnn), _ptr(aad), m.aa_stride, n, _ptr(mk2), None, None, _stream()), "tc")
    torch.cuda.synchronize()
    v1, p1 = _decode(mk1, n, nets, J2)
    v2, p2 = _decode(mk2, n, nets, J2)
    # bf16x3 products carry ~2^-16 relative error of the TERMS; channel maxima are sums with cancellation,
    # so the error is judged against the scale of the layer (max |r2|), not against each small value
    err = np.max(np.abs(v1 - v2)) / np.abs(v1).max()
    assert err < 2e-5, f"max err of channel maxima relative to the layer scale {err:.3e}"
    diff = p1 != p2
    if diff.any():                       # arg-max may only differ where two positions are within rounding
        en = port.PortEnergy(w)
        x = port.aa_to_onehot(aa)
        for (b, k, j) in zip(*np.nonzero(diff)):
            net = en.cnn[k]
            z = torch.relu(torch.nn.functional.conv1d(x[b:b + 1].transpose(1, 2), net["W0"], net["b0"]).transpose(1, 2))
            r2 = torch.relu(torch.nn.functional.linear(z, net["W1"], net["b1"]))[0, :, j].numpy()
            assert abs(r2[p1[b, k, j]] - r2[p2[b, k, j]]) <= 2e-5 * max(abs(r2).max(), 1e-6), "arg-max differs beyond a near tie"
    # fitness through both paths vs the oracle
    en = port.PortEnergy(w)
    with torch.no_grad():
        fit_ref = en.fitness(port.aa_to_onehot(aa)).numpy()
    for impl in ("simt", "tc"):
        m.cnn_forward_impl = impl
        E, fit, G, Ep = m.energy(aad)
        torch.cuda.synchronize()
        e = np.max(np.abs(fit.cpu().numpy() - fit_ref) / np.maximum(np.abs(fit_ref), 1e-2))
        assert e < 1e-4, f"{impl}: fitness rel err {e:.3e}"


@pytest.mark.parametrize("L,n", [(40, 9), (104, 40), (238, 150)])
def test_tc_backward_matches_simt_and_oracle(L, n):
    from ppde_b200.engine import PoEModel
    w = port.synthetic_weights(L, seed=L + 1, lamda=3.0, window=(1, L - 2))
    m = PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    rng = np.random.default_rng(L)
    aa = np.tile(w.wt, (n, 1))
    for b in range(1, n):
        pos = rng.integers(0, L, size=min(b, L))
        aa[b, pos] = rng.integers(0, 20, size=pos.shape[0])
    aa[n - 1, 5:11] = aa[n - 1, 20:26]             # repeated 5-mers: exact max-pool ties
    pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = aa
    aad = torch.from_numpy(pad).to(m.device)
    out = {}
    for impl in ("simt", "tc"):
        m.cnn_forward_impl = impl
        m.cnn_backward_impl = impl
        E, fit, G, Ep = m.energy(aad)
        torch.cuda.synchronize()
        out[impl] = (E.cpu().numpy(), fit.cpu().numpy(), G.cpu().numpy())
    gs, gt = out["simt"][2], out["tc"][2]
    scale = np.abs(gs).max(axis=(1, 2), keepdims=True)
    err = np.max(np.abs(gs - gt) / scale)
    assert err < 2e-5, f"tc vs simt gradient, relative to each chain's max |g|: {err:.3e}"
    en = port.PortEnergy(w)
    k = min(n, 12)
    e_ref, f_ref, g_ref = en.get_energy_and_grads(port.aa_to_onehot(aa[-k:]))
    g_ref = g_ref.numpy()
    gscale = np.maximum(np.abs(g_ref), np.abs(g_ref).max(axis=(1, 2), keepdims=True) * 1e-2)
    err = np.max(np.abs(gt[-k:] - g_ref) / gscale)
    assert err < 1e-4, f"tc gradient vs oracle autograd: {err:.3e}"
    e_ref = e_ref.numpy()
    err = np.max(np.abs(out["tc"][0][-k:] - e_ref) / np.maximum(np.abs(e_ref), abs(m.wt_H)))
    assert err < 1e-4, f"tc energy vs oracle: {err:.3e}"
