"""The mathematics behind the incremental CNN forward and the delta backward, checked on the CPU against the oracle
port's own autograd (no CUDA involved):

  * the max-pool over positions equals the maximum over 16-position blocks of (block maximum, first arg-max), and a block
    can only change if it holds a conv row p in [i-4, i] of a changed residue i        (cnn_dirty_kernel, cnn_forward_inc_kernel)
  * the adjoint rows A[p,:] = relu'(r1[p,:]) . sum_{j: argmax_j = p, m_j > 0} d_j W1[j,:] of two states differ only on
    D = D0 U {both ends of every moved winner}, D0 = U_{i changed} [i-4, i], and the signed entry list of
    cnn_winner_delta_kernel reproduces the difference row by row                        (cnn_backward_delta_kernel)
  * d fit_k / dx = col2im(W0^T A), so the gradient of the proposal is the gradient of the current state plus the
    col2im of the changed rows                                                           (cnn_grad_combine_delta_kernel)
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import ppde_port as port


def _layers(net, aa):
    """conv pre-activation mask, max-pool values / first arg-max, adjoint rows of ONE state (float64 for exact comparisons)."""
    x = port.aa_to_onehot(aa[None]).double()
    W0, b0, W1, b1, d = (net[k].double() for k in ("W0", "b0", "W1", "b1", "d"))
    r1 = F.relu(F.conv1d(x.transpose(1, 2), W0, b0).transpose(1, 2))[0]          # [P, C]
    r2 = F.relu(F.linear(r1, W1, b1))                                            # [P, 2C]
    m, arg = torch.max(r2, dim=0)                                                # first arg-max on ties (CPU torch.max)
    mask = (r1 > 0).double()
    P = r1.shape[0]
    A = torch.zeros(P, r1.shape[1], dtype=torch.float64)
    for j in range(r2.shape[1]):
        if m[j] > 0:
            A[arg[j]] += d[j] * W1[j]
    return r2, m, arg, mask, A * mask, W0, W1, d


def _col2im(A, W0, L):
    """d fit / dx [L, 20] from the adjoint rows: G[i, a] = sum_t sum_c W0[c, a, t] A[i - t, c]"""
    P = A.shape[0]
    G = torch.zeros(L, 20, dtype=torch.float64)
    for t in range(5):
        G[t:t + P] += A @ W0[:, :, t]                                            # [P, 20]
    return G


def test_block_maxima_and_dirty_blocks():
    L = 60
    w = port.synthetic_weights(L, seed=3, lamda=1.0)
    en = port.PortEnergy(w)
    rng = np.random.default_rng(0)
    P = L - 4
    NB = (P + 15) // 16
    for trial in range(6):
        x = w.wt.copy()
        x[rng.integers(0, L, size=6)] = rng.integers(0, 20, size=6)
        y = x.copy()
        pos = rng.integers(0, L, size=int(rng.integers(1, 4)))
        y[pos] = (y[pos] + rng.integers(1, 20, size=pos.shape[0])) % 20
        dirty = set()
        for i in np.nonzero(x != y)[0]:
            for p in range(max(i - 4, 0), min(i, P - 1) + 1):
                dirty.add(p >> 4)
        for net in en.cnn:
            r2x, mx, ax, *_ = _layers(net, x)
            r2y, my, ay, *_ = _layers(net, y)
            for q in range(NB):
                bx, by = r2x[16 * q:16 * q + 16], r2y[16 * q:16 * q + 16]
                if q not in dirty:
                    assert torch.equal(bx, by), "a block without a touched conv row changed"
            # chain-level (max, first arg-max) from block-level ones, ties to the lowest position
            for r2, m, a in ((r2x, mx, ax), (r2y, my, ay)):
                best_v = torch.full_like(m, -1.0); best_p = torch.zeros_like(a)
                for q in range(NB):
                    v, i = torch.max(r2[16 * q:16 * q + 16], dim=0)
                    take = v > best_v                                           # strict: earlier blocks win ties
                    best_v = torch.where(take, v, best_v); best_p = torch.where(take, i + 16 * q, best_p)
                assert torch.equal(best_v, m) and torch.equal(best_p, a)


def test_adjoint_rows_change_only_on_touched_positions_and_delta_gradient():
    L = 60
    w = port.synthetic_weights(L, seed=5, lamda=2.0)
    en = port.PortEnergy(w)
    rng = np.random.default_rng(1)
    P = L - 4
    for trial in range(6):
        x = w.wt.copy()
        x[rng.integers(0, L, size=8)] = rng.integers(0, 20, size=8)
        y = x.copy()
        pos = rng.integers(0, L, size=int(rng.integers(1, 4)))
        y[pos] = (y[pos] + rng.integers(1, 20, size=pos.shape[0])) % 20
        D0 = set()
        for i in np.nonzero(x != y)[0]:
            D0.update(range(max(i - 4, 0), min(i, P - 1) + 1))
        g_delta = torch.zeros(L, 20, dtype=torch.float64)
        for net in en.cnn:
            _, mx, ax, maskx, Ax, W0, W1, d = _layers(net, x)
            _, my, ay, masky, Ay, *_ = _layers(net, y)
            px = torch.where(mx > 0, ax, torch.full_like(ax, -1))              # -1 = dead channel (relu'(0) = 0)
            py = torch.where(my > 0, ay, torch.full_like(ay, -1))
            moved = px != py
            touched = set(D0)
            for j in torch.nonzero(moved).flatten().tolist():
                touched.update(p for p in (int(px[j]), int(py[j])) if p >= 0)
            for p in range(P):
                if p not in touched:
                    assert torch.equal(Ax[p], Ay[p]), "an adjoint row outside the touched set changed"
            # the signed entry list of cnn_winner_delta_kernel: every winner sitting on a row of D0 (both sides) and both ends
            # of every moved winner; the masks of the entry's own side
            dA = torch.zeros_like(Ax)
            for j in range(W1.shape[0]):
                for side, pj, mask, sign in ((0, int(py[j]), masky, 1.0), (1, int(px[j]), maskx, -1.0)):
                    if pj >= 0 and (bool(moved[j]) or pj in D0):
                        dA[pj] += sign * d[j] * W1[j] * mask[pj]
            assert torch.allclose(dA, Ay - Ax, rtol=0, atol=1e-12), "entry list does not reproduce the change of the adjoint rows"
            g_delta += _col2im(dA, W0, L)
        # gradient of the CNN part: autograd at y == autograd at x + lamda / n_nets * col2im(change)
        def cnn_grad(aa):
            xx = port.aa_to_onehot(aa[None]).double().requires_grad_()
            nets64 = [{k: v.double() for k, v in net.items()} for net in en.cnn]
            fit = torch.mean(torch.stack([en._cnn_one(net, xx) for net in nets64], 0), 0).squeeze()
            return torch.autograd.grad([fit], xx)[0][0]
        gx, gy = cnn_grad(x), cnn_grad(y)
        assert torch.allclose(gx + g_delta / len(en.cnn), gy, rtol=0, atol=1e-10)
