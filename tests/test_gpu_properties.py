"""Size-independent properties of the CUDA path at sizes the oracle cannot reach, plus sharding invariance."""
import numpy as np
import pytest
import torch

from oracle import ppde_port as port

pytestmark = pytest.mark.gpu


def _setup(L, lamda, window=None, seed=0):
    from ppde_b200.engine import PoEModel
    w = port.synthetic_weights(L, seed=seed, lamda=lamda, window=window)
    m = PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    return w, m


def _wt_pop(m, w, n):
    pad = np.zeros((n, m.aa_stride), dtype=np.uint8)
    pad[:, :w.L] = w.wt
    return torch.from_numpy(pad).to(m.device)


@pytest.mark.parametrize("L,n,nmut,paper,pas", [(238, 4096, 0, False, 2), (104, 2048, 10, False, 2), (96, 1024, 5, True, 3)])
def test_cached_state_equals_fresh_evaluation(L, n, nmut, paper, pas):
    """After T iterations of incremental updates, the cached (E, fit, G) of every chain equals a from-scratch
    evaluation of its current sequence (bounds the drift of the incremental Potts field), and the edit-distance
    constraint of the hard mode holds for every chain."""
    from ppde_b200.engine import ChainEngine
    w, m = _setup(L, 2.0, window=(2, L - 3))
    T = 12
    eng = ChainEngine(m, n, pas, nmut, paper, seed=3, num_steps=T)
    eng.init_population(_wt_pop(m, w, n))
    eng.run_steps(T, use_graph=True)
    torch.cuda.synchronize()
    E, fit, G, Ep = m.energy(eng.aa)
    rows = eng.row_cur.long()
    Gc = eng.G[rows].view(n, L, 20)
    scale = torch.clamp(E.abs(), min=abs(m.wt_H))
    assert float(((eng.E - E).abs() / scale).max()) < 1e-4
    assert float(((eng.fit - fit).abs() / torch.clamp(fit.abs(), min=1e-2)).max()) < 1e-4
    gs = G.abs().amax(dim=(1, 2), keepdim=True)
    assert float(((Gc - G).abs() / gs).max()) < 1e-4
    dist, _ = eng.population_metrics()
    if nmut and not paper:
        assert int(dist.max()) < nmut            # chains at/over the threshold were reset to WT (ppde.py:148-153)
    # history row T is the energy recorded at the last iteration; best-of-history is its running maximum
    assert torch.equal(eng.best_E, eng.E_hist.max(dim=0).values)
    acc = eng.accept.bool()
    assert 0 < int(acc.sum()) <= n


def test_results_do_not_depend_on_sharding():
    """Streams are indexed by global chain id: one engine of 96 chains == three engines of 32 chains."""
    from ppde_b200.engine import ChainEngine
    L, n, T = 60, 96, 6
    w, m = _setup(L, 1.5)
    whole = ChainEngine(m, n, 2, 4, False, seed=11, num_steps=T)
    whole.init_population(_wt_pop(m, w, n))
    whole.run_steps(T, use_graph=False)
    parts = []
    for r in range(3):
        e = ChainEngine(m, 32, 2, 4, False, seed=11, chain_offset=32 * r, num_steps=T)
        e.init_population(_wt_pop(m, w, 32))
        e.run_steps(T, use_graph=True)             # graph replay on the shards, eager on the whole: same numbers
        parts.append(e)
    torch.cuda.synchronize()
    assert torch.equal(whole.E_hist, torch.cat([p.E_hist for p in parts], dim=1))
    assert torch.equal(whole.aa, torch.cat([p.aa for p in parts], dim=0))
    assert torch.equal(whole.best_aa, torch.cat([p.best_aa for p in parts], dim=0))
    assert torch.equal(whole.idx, torch.cat([p.idx for p in parts], dim=1))


def test_materialised_uniforms_equal_philox_stream():
    """Parity mode (uniforms handed in as a buffer) and the in-kernel Philox stream give the same chain."""
    from ppde_b200 import philox
    from ppde_b200.engine import ChainEngine
    L, n, T = 40, 16, 5
    w, m = _setup(L, 1.0)
    a = ChainEngine(m, n, 2, 0, False, seed=21, num_steps=T)
    b = ChainEngine(m, n, 2, 0, False, seed=21, num_steps=T)
    a.init_population(_wt_pop(m, w, n)); b.init_population(_wt_pop(m, w, n))
    for t in range(T):
        u = np.stack([philox.proposal_uniforms(21, t, s, np.arange(n), 20 * L) for s in range(3)])
        a.step()
        b.step(uniforms=torch.from_numpy(u).to(m.device))
    torch.cuda.synchronize()
    assert torch.equal(a.E_hist, b.E_hist) and torch.equal(a.aa, b.aa)


def test_sampler_api_six_tuple_and_supervised_energy():
    """PPDE_PAS.run through the reference-shaped API (6-tuple contract, ppde.py:191-192) and the CNN-only energy."""
    import argparse
    from ppde_b200.energy import ProteinProductOfExperts, ProteinSupervised
    from ppde_b200.sampler import PPDE_PAS
    L, n, T = 50, 24, 7
    w = port.synthetic_weights(L, seed=2, lamda=3.0, window=(1, L - 2))
    en = ProteinProductOfExperts.from_arrays(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    args = argparse.Namespace(ppde_pas_length=2, nmut_threshold=3, paper_results=False, seed=4, ppde_verbose=False)
    pop = en.wt_onehot.repeat(n, 1, 1)
    np.random.seed(0)
    calls = []

    def oracle(x):
        calls.append(tuple(x.shape))
        return x.sum((1, 2))
    out = PPDE_PAS(args).run(pop, T, en, 1, L - 2, oracle, log_every=3)
    best_x, best_e, best_f, e_hist, f_hist, traj = out
    assert tuple(best_x.shape) == (n, L, 20) and best_x.device == pop.device
    assert best_e.shape == (n,) and best_f.shape == (n,) and e_hist.shape == (T + 1, n) and f_hist.shape == (T + 1, n)
    assert len(traj) == T + 1 and traj[0].shape == (L, 20)
    assert np.array_equal(best_e, e_hist.max(0))
    assert calls == [(n, L, 20)] * 3                      # t=0 and iterations 2, 5 (i>0 and (i+1)%3==0)
    # the port run on the same streams gives the same histories
    pe = port.PortEnergy(w)
    ref = port.PortSampler(2, 3, False, seed=4).run(pe.wt_onehot.repeat(n, 1, 1), T, pe)
    assert np.max(np.abs(e_hist - ref[3]) / np.maximum(np.abs(ref[3]), abs(en.model.wt_H))) < 1e-4
    assert np.array_equal(best_x.argmax(-1).cpu().numpy(), ref[0].argmax(-1).numpy())
    # get_energy / get_energy_and_grads on a one-hot batch (reference signature)
    x = port.aa_to_onehot(np.random.default_rng(0).integers(0, 20, size=(5, L))).to("cuda:0")
    e, f, g = en.get_energy_and_grads(x)
    e2, f2, g2 = pe.get_energy_and_grads(x.cpu())
    assert np.max(np.abs(g.cpu().numpy() - g2.numpy())) / np.abs(g2.numpy()).max() < 1e-4
    # CNN-only energy (ProteinSupervised, energy.py:143-164)
    sup = ProteinSupervised.from_arrays(w.wt, w.cnn, device="cuda:0")
    fs, fs2, gs = sup.get_energy_and_grads(x)
    xr = x.cpu().requires_grad_()
    fr = pe.fitness(xr)
    gr = torch.autograd.grad([fr.sum()], xr)[0]
    assert torch.equal(fs, fs2)
    assert np.max(np.abs(fs.cpu().numpy() - fr.detach().numpy()) / np.maximum(np.abs(fr.detach().numpy()), 1e-2)) < 1e-4
    assert np.max(np.abs(gs.cpu().numpy() - gr.numpy())) / np.abs(gr.numpy()).max() < 1e-4
    out = PPDE_PAS(args).run(pop, 3, sup, 0, L - 1, None, log_every=50)
    assert out[3].shape == (4, n)
