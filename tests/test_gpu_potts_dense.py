"""Dense tcgen05 Potts contraction vs the row-gather kernel and the oracle (fp32 einsum + autograd).

Reference: PottsModel.hamiltonian / forward(delta=True), ppde/nets.py:282-299; gradient ppde/energy.py:106-108."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _states(rng, wt, n, nmut):
    aa = np.tile(wt, (n, 1)).astype(np.uint8)
    for b in range(n):
        pos = rng.integers(0, wt.shape[0], size=nmut)
        aa[b, pos] = rng.integers(0, 20, size=nmut)
    return aa


@pytest.mark.parametrize("L,window,n", [(40, (2, 36), 37), (96, None, 300), (104, (22, 97), 513), (238, None, 700)])
def test_dense_matches_gather_and_oracle(L, window, n):
    from oracle import ppde_port as port
    from ppde_b200 import _lib
    from ppde_b200.engine import PoEModel, _ptr, _stream

    w = port.synthetic_weights(L, seed=L + 1, lamda=1.0, window=window)
    m = PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    rng = np.random.default_rng(L)
    aa = _states(rng, w.wt, n, 12)
    pad = np.zeros((n, m.aa_stride), dtype=np.uint8)
    pad[:, :L] = aa
    aad = torch.from_numpy(pad).to(m.device)
    out = {}
    for impl in ("gather", "dense"):
        Gp = torch.full((n, m.D), float("nan"), dtype=torch.float32, device=m.device)
        Ep = torch.empty(n, dtype=torch.float32, device=m.device)
        m.potts_full(aad, n, _ptr(Gp), _ptr(Ep), _stream(), impl=impl)
        torch.cuda.synchronize()
        out[impl] = (Gp.cpu().numpy(), Ep.cpu().numpy())
    gg, eg = out["gather"]
    gd, ed = out["dense"]
    assert np.isfinite(gd).all()
    gscale = np.abs(gg).max()
    # two fp32 summation orders of Lp terms (tensor-core K order vs position order): ~1e-6 of the largest entry
    assert np.abs(gd - gg).max() <= 1e-5 * gscale, np.abs(gd - gg).max() / gscale
    escale = max(np.abs(eg).max(), abs(m.wt_H))
    assert np.abs(ed - eg).max() <= 1e-5 * escale
    # oracle: fp32 einsum Hamiltonian + autograd on a sample of the chains
    en = port.PortEnergy(w)
    k = min(n, 24)
    x = port.aa_to_onehot(aa[:k]).requires_grad_(True)
    e_ref = en.potts_delta(x) if hasattr(en, "potts_delta") else None
    if e_ref is not None:
        g_ref = torch.autograd.grad(e_ref.sum(), x)[0].numpy()[:, m.win_lo:m.win_lo + m.Lp, :].reshape(k, -1)
        assert np.abs(gd[:k] - g_ref).max() <= 1e-4 * max(np.abs(g_ref).max(), 1e-6)
        assert np.abs(ed[:k] - e_ref.detach().numpy()).max() <= 1e-4 * escale
