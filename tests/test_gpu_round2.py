"""GPU parity, round 2: the exact-refresh path of the delta backward, the direct known-answer tests of the proposal arithmetic
and of the no-grad energy, the device-side log_every report, engine isolation and the host-buffer API forms."""
import argparse
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import ppde_port as port

pytestmark = pytest.mark.gpu
GOLD = port.GOLDEN_DIR


def _meta(z):
    return dict(potts_seed=int(z["potts_seed"]), sigma_j=float(z["sigma_j"]), sigma_h=float(z["sigma_h"]),
                symmetric=bool(z["symmetric"]), zero_diag=bool(z["zero_diag"]))


def _model(w):
    from ppde_b200.engine import PoEModel
    return PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")


def _wt_pop(m, w, n):
    pad = np.zeros((n, m.aa_stride), dtype=np.uint8)
    pad[:, :w.L] = w.wt
    return torch.from_numpy(pad).to(m.device)


def _rel(a, b, floor):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


# ------------------------------------------------------------------------------------------- refresh path
def test_refresh_path_vs_port_pabp():
    """T = 70 graph-replayed iterations with the shipped PABP CNNs: crosses the exact-refresh iterations t = 31 and t = 63 (second
    captured graph) and the hand-over back to the delta backward.  The whole run - every history row, the best-of-history
    sequences, the tracked trajectory - must equal the oracle port's (the port is pinned to the unmodified reference), and the
    cached gradient rows after 70 updates must equal a fresh evaluation."""
    from ppde_b200.engine import ChainEngine
    z = np.load(os.path.join(GOLD, "traj_pabp_hard.npz"))
    w = port.golden_weights(str(z["prot"]), z["window"], float(z["lamda"]), **_meta(z))
    m = _model(w)
    assert m.cnn_bwd_delta and m.bwd_refresh == 32
    n, T, seed = 8, 70, 17
    eng = ChainEngine(m, n, 2, 0, False, seed=seed, num_steps=T, traj_chain=3)
    eng.init_population(_wt_pop(m, w, n))
    eng.run_steps(T, use_graph=True)
    torch.cuda.synchronize()
    assert set(eng._graph) == {True, False}, "both graph variants (exact refresh and delta) must have been replayed"
    en = port.PortEnergy(w)
    ref = port.PortSampler(2, 0, False, seed=seed).run(en.wt_onehot.repeat(n, 1, 1), T, en, random_idx=3)
    assert _rel(eng.E_hist.cpu().numpy(), ref[3], abs(m.wt_H)) < 1e-4
    assert _rel(eng.fit_hist.cpu().numpy(), ref[4], 1e-2) < 1e-4
    assert np.array_equal(eng.best_aa.cpu().numpy()[:, :w.L], ref[0].argmax(-1).numpy())
    assert np.array_equal(eng.traj_aa.cpu().numpy()[:, :w.L], np.stack([t.argmax(-1) for t in ref[5]]))
    E, fit, G, _ = m.energy(eng.aa)
    Gc = eng.G[eng.row_cur.long()].view(n, w.L, 20)
    gs = G.abs().amax(dim=(1, 2), keepdim=True)
    assert float(((Gc - G).abs() / gs).max()) < 1e-5
    with pytest.raises(ValueError):
        eng.run_steps(1)                       # the engine was allocated for T iterations: no room for another history row


def test_refresh_path_delta_vs_exact_backward_gfp_length():
    """L = 238 synthetic, 64 chains, 70 iterations: the default engine (delta backward + exact refresh every 32nd
    iteration, CUDA graphs) against an engine that runs the exact backward every iteration.  Proposal indices, accept
    decisions and states must agree bit for bit at every iteration - in particular right after the refresh iterations, whose
    proposal rows come from the same exact kernels.  The two gradients differ by ~1e-6 of max|G| between refreshes, so a
    proposal race may flip at a near tie (north star: "identical except at stated near-tie decisions"): a chain that flips is
    reported and leaves the comparison; at most 2 of the 64 may.  The first 16 chains are also checked against the port."""
    from ppde_b200.engine import ChainEngine
    L, n, T, seed = 238, 64, 70, 5
    w = port.synthetic_weights(L, seed=1, lamda=15.0)
    m = _model(w)
    assert m.cnn_bwd_delta
    a = ChainEngine(m, n, 2, 0, False, seed=seed, num_steps=T)
    a.init_population(_wt_pop(m, w, n))
    m.cnn_bwd_delta = False
    try:
        b = ChainEngine(m, n, 2, 0, False, seed=seed, num_steps=T)
        b.init_population(_wt_pop(m, w, n))
    finally:
        m.cnn_bwd_delta = True
    assert a.delta and not b.delta
    same = torch.ones(n, dtype=torch.bool, device=m.device)        # chains that have not flipped at a near tie
    refreshes = 0
    for t in range(T):
        a.run_steps(1, use_graph=True)
        b.run_steps(1, use_graph=True)
        torch.cuda.synchronize()
        same &= (a.idx == b.idx).all(dim=0) & (a.accept == b.accept) & (a.aa == b.aa).all(dim=1)
        if a.full_backward_at(t):              # refresh iteration: the proposal rows were produced by the same exact kernels
            refreshes += 1
            ra, rb = a.rows_y.long()[same], b.rows_y.long()[same]
            assert torch.equal(ra, rb)
            assert torch.equal(a.G[ra], b.G[rb]), f"refresh iteration t={t}: rows differ from the exact run"
    assert refreshes == 2
    flipped = int((~same).sum())
    assert flipped <= 2, f"{flipped} of {n} chains diverged between the delta and the exact backward"
    assert torch.equal(a.E_hist[:, same], b.E_hist[:, same])
    ga, gb = a.G[a.row_cur.long()][same], b.G[b.row_cur.long()][same]
    assert float(((ga - gb).abs() / gb.abs().amax(dim=1, keepdim=True)).max()) < 1e-5
    k = 16
    en = port.PortEnergy(w)
    ref = port.PortSampler(2, 0, False, seed=seed).run(en.wt_onehot.repeat(k, 1, 1), T, en)
    ok = same[:k].cpu().numpy()
    assert ok.sum() >= k - 2
    assert _rel(b.E_hist.cpu().numpy()[:, :k], ref[3], abs(m.wt_H)) < 1e-4          # the exact-backward engine: every chain
    assert _rel(a.E_hist.cpu().numpy()[:, :k][:, ok], ref[3][:, ok], abs(m.wt_H)) < 1e-4
    assert np.array_equal(a.best_aa.cpu().numpy()[:k, :L][ok], ref[0].argmax(-1).numpy()[ok])


def test_sample_of_a_large_population_vs_port():
    """4,096 chains at L = 238 (the bench configuration's shape): 16 of them, taken from the middle of the population, are
    re-run by the oracle port on the same streams (chain_offset) and must agree."""
    from ppde_b200.engine import ChainEngine
    L, n, T, seed, lo, k = 238, 4096, 10, 9, 2040, 16
    w = port.synthetic_weights(L, seed=0, lamda=15.0)
    m = _model(w)
    eng = ChainEngine(m, n, 2, 0, False, seed=seed, num_steps=T)
    eng.init_population(_wt_pop(m, w, n))
    eng.run_steps(T, use_graph=True)
    torch.cuda.synchronize()
    en = port.PortEnergy(w)
    ref = port.PortSampler(2, 0, False, seed=seed, chain_offset=lo).run(en.wt_onehot.repeat(k, 1, 1), T, en)
    assert _rel(eng.E_hist.cpu().numpy()[:, lo:lo + k], ref[3], abs(m.wt_H)) < 1e-4
    assert np.array_equal(eng.best_aa.cpu().numpy()[lo:lo + k, :L], ref[0].argmax(-1).numpy())


# ------------------------------------------------------------------------------------------- direct KATs
@pytest.mark.parametrize("name", ["pabp", "pabp_asym", "ube4b", "gfp"])
def test_get_energy_nograd_vs_reference(name):
    """`get_energy` (ppde/energy.py:97-101) - the no-grad launcher branch - against the reference's own outputs."""
    from ppde_b200.energy import ProteinProductOfExperts
    z = np.load(os.path.join(GOLD, f"kat_energy_{name}.npz"))
    w = port.golden_weights(str(z["prot"]), z["window"], float(z["lamda"]), **_meta(z))
    en = ProteinProductOfExperts.from_arrays(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    x = port.aa_to_onehot(z["aa"]).to("cuda:0")
    e, fit = en.get_energy(x)
    torch.cuda.synchronize()
    assert _rel(e.cpu().numpy(), z["e_nograd"], abs(float(z["wt_H"]))) < 1e-4
    assert _rel(fit.cpu().numpy(), z["fit_nograd"], 1e-2) < 1e-4
    assert _rel(en.get_unsupervised_expert(x).cpu().numpy(), z["potts_delta"], abs(float(z["wt_H"]))) < 1e-4


def test_proposal_arithmetic_kat_vs_reference():
    """mut_distance / mutation_mask / safe_logits_to_probs + Categorical (ppde/utils.py:5-28,106-111) through the device
    functions of the proposal kernels (ppde_pas_kat) against tests/golden/kat_int.npz, made by the unmodified reference:
    integers bit-exact, probabilities and the log-probability to 1e-4 (measured ~3e-7: ex2.approx + one reciprocal)."""
    from ppde_b200 import _lib
    lib = _lib.load()
    z = np.load(os.path.join(GOLD, "kat_int.npz"))
    n, L = z["aa"].shape
    dev = "cuda:0"
    stride = (L + 15) // 16 * 16
    pad = np.zeros((n, stride), dtype=np.uint8); pad[:, :L] = z["aa"]
    aa = torch.from_numpy(pad).to(dev)
    wt = torch.from_numpy(z["wt"]).to(dev)
    logits = torch.from_numpy(z["logits"]).to(dev).contiguous()
    idx = torch.from_numpy(z["idx"].astype(np.int32)).to(dev)
    dist = torch.zeros(n, dtype=torch.int32, device=dev)
    mask = torch.zeros(n, 20 * L, dtype=torch.uint8, device=dev)
    probs = torch.zeros(n, 20 * L, dtype=torch.float32, device=dev)
    logp = torch.zeros(n, dtype=torch.float32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(lib.ppde_pas_kat(p(aa), stride, p(wt), n, L, p(logits), p(idx), p(dist), p(mask), p(probs), p(logp),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pas_kat")
    torch.cuda.synchronize()
    assert np.array_equal(dist.cpu().numpy(), z["dist"].astype(np.int32))
    assert np.array_equal(mask.cpu().numpy().reshape(n, L, 20).astype(bool), z["mask"])
    assert _rel(probs.cpu().numpy(), z["probs_norm"], 1e-30) < 1e-4
    assert _rel(logp.cpu().numpy(), z["log_prob"], 1.0) < 1e-4
    # the population-metrics kernel computes the same distances
    d2 = torch.zeros(n, dtype=torch.int32, device=dev)
    _lib.check(lib.ppde_population_metrics(p(aa), stride, n, L, p(wt), p(d2), C.c_void_p(0),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)), "metrics")
    assert np.array_equal(d2.cpu().numpy(), z["dist"].astype(np.int32))


# ------------------------------------------------------------------------------------------- log_every report
@pytest.mark.parametrize("n,L", [(1, 31), (37, 96), (5000, 238)])
def test_population_report_vs_numpy(n, L):
    """Device-side report (radix-select quantiles, exact unique count, integer sums, top-k) against numpy / the port's
    restatement of scripts/make_figures.py:29-49 and torch.topk."""
    from ppde_b200 import _lib
    from ppde_b200 import dist as D
    lib = _lib.load()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(n)
    wt = rng.integers(0, 20, L).astype(np.uint8)
    aa = np.tile(wt, (n, 1))
    for b in range(n):                              # many duplicates (b % 7 distinct mutation patterns + exact copies of WT)
        r2 = np.random.default_rng(b % 7 if b % 3 else 1000 + b)
        pos = r2.integers(0, L, size=b % 5)
        aa[b, pos] = r2.integers(0, 20, size=pos.shape[0])
    stride = (L + 15) // 16 * 16
    pad = np.zeros((n, stride), dtype=np.uint8); pad[:, :L] = aa
    e = rng.standard_normal(n).astype(np.float32) * 3
    e[rng.integers(0, n, size=max(n // 5, 1))] = 1.25          # ties, also at the top-k boundary
    if n > 20:
        e[:12] = e.max()
    f = rng.standard_normal(n).astype(np.float32)
    g = np.abs(rng.standard_normal(n)).astype(np.float32)
    acc = (rng.random(n) < 0.9).astype(np.uint8)
    rep = D.PopulationReporter(lib, dev, n, L, stride, torch.from_numpy(wt).to(dev), top_k=16).report(
        torch.from_numpy(e).to(dev), torch.from_numpy(f).to(dev), torch.from_numpy(g).to(dev), torch.from_numpy(acc).to(dev),
        torch.from_numpy(pad).to(dev), chain_lo=100)
    for key, v in (("energy_q", e), ("fitness_q", f), ("oracle_q", g)):
        want = np.quantile(v, [0.5, 0.9])
        assert np.allclose(rep[key], want, rtol=1e-6, atol=1e-7), (key, rep[key], want)
    assert rep["accepted"] == float(acc.sum())
    mean, std = port.n_hops(aa, wt)
    assert rep["mean_dist"] == pytest.approx(mean, rel=1e-12) and rep["std_dist"] == pytest.approx(std, rel=1e-9, abs=1e-12)
    assert rep["diversity_pct"] == pytest.approx(port.diversity_percent(aa), rel=1e-12)
    assert rep["unique"] == len({bytes(r) for r in aa})
    k = min(16, n)
    tv, ti = torch.topk(torch.from_numpy(e), k)
    assert np.array_equal(rep["topk_energy"], tv.numpy())
    # ties -> lowest chain id first (the order np.argsort(kind='stable') of -e gives)
    order = np.argsort(-e.astype(np.float64), kind="stable")[:k]
    assert np.array_equal(rep["topk_chain"], order + 100)
    assert np.array_equal(rep["topk_aa"], aa[order])


def test_sampler_logging_matches_host_restatement(capsys):
    """PPDE_PAS.run with log_every: the printed report lines (format of ppde.py:164-168) carry the numbers a host-side
    numpy restatement computes from the engine's own state."""
    from ppde_b200.energy import ProteinProductOfExperts
    from ppde_b200.ridge import AugmentedLinearRegression
    from ppde_b200.sampler import PPDE_PAS
    L, n, T = 50, 64, 6
    w = port.synthetic_weights(L, seed=4, lamda=3.0, window=(1, L - 2))
    en = ProteinProductOfExperts.from_arrays(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    rng = np.random.default_rng(0)
    orc = AugmentedLinearRegression(en.model, rng.standard_normal((4, 1 + 20 * L)).astype(np.float32) * 0.1,
                                    rng.standard_normal(4).astype(np.float32), np.ones(4))
    args = argparse.Namespace(ppde_pas_length=2, nmut_threshold=4, paper_results=False, seed=2, ppde_verbose=True)
    smp = PPDE_PAS(args)
    np.random.seed(1)
    out = smp.run(en.wt_onehot.repeat(n, 1, 1), T, en, 1, L - 2, orc, log_every=3)
    printed = capsys.readouterr().out
    assert [i for i, _ in smp.reports] == [2, 5]
    eng = smp.engine
    i, rep = smp.reports[-1]
    e_row, f_row = out[3][i + 1], out[4][i + 1]
    assert np.allclose(rep["energy_q"], np.quantile(e_row, [0.5, 0.9]), rtol=1e-6)
    assert np.allclose(rep["fitness_q"], np.quantile(f_row, [0.5, 0.9]), rtol=1e-6)
    aa = eng.aa.cpu().numpy()[:, :L]
    assert rep["accepted"] == float(eng.accept.sum().item())
    assert rep["mean_dist"] == pytest.approx(port.n_hops(aa, w.wt)[0])
    assert rep["diversity_pct"] == pytest.approx(port.diversity_percent(aa))
    gt = orc.score_engine(eng).cpu().numpy()
    assert np.allclose(rep["oracle_q"], np.quantile(gt, [0.5, 0.9]), rtol=1e-6)
    assert f"[Iteration {i}] energy: 50% {rep['energy_q'][0]:.3f}, 90% {rep['energy_q'][1]:.3f}" in printed
    assert f"   # accepted = {rep['accepted']}" in printed and "[Iteration 0] oracle fit 50%" in printed
    assert np.array_equal(rep["topk_energy"], np.sort(eng.E_hist[i + 1].cpu().numpy())[::-1][:16])


# ------------------------------------------------------------------------------------------- engine isolation, API forms
def test_two_engines_of_different_size_do_not_share_scratch():
    """Each ChainEngine owns its kernel scratch (the pointers live in its captured graphs): interleaving two engines of
    different sizes on one model gives exactly what each gives alone."""
    from ppde_b200.engine import ChainEngine
    L, T = 96, 8
    w = port.synthetic_weights(L, seed=6, lamda=2.0)
    m = _model(w)

    def fresh(n, seed):
        e = ChainEngine(m, n, 2, 5, False, seed=seed, num_steps=T)
        e.init_population(_wt_pop(m, w, n))
        return e
    a0, b0 = fresh(40, 1), fresh(300, 2)
    a0.run_steps(T); b0.run_steps(T)
    a1, b1 = fresh(40, 1), fresh(300, 2)
    for _ in range(T // 2):
        a1.run_steps(2); b1.run_steps(2)
        m.energy(b1.aa)                                  # a stand-alone evaluation in between (the model's own scratch)
    torch.cuda.synchronize()
    assert torch.equal(a0.E_hist, a1.E_hist) and torch.equal(a0.aa, a1.aa)
    assert torch.equal(b0.E_hist, b1.E_hist) and torch.equal(b0.aa, b1.aa)


def test_host_population_forms_give_the_same_run():
    """The population may cross the API as a device one-hot (reference form), a HOST one-hot (reduced to residues by the host
    cores before the copy) or uint8 residues (args.ppde_residue_io): same chains, and best_x comes back in the form and on
    the device it came in."""
    from ppde_b200.energy import ProteinProductOfExperts
    from ppde_b200.sampler import PPDE_PAS
    L, n, T = 60, 48, 5
    w = port.synthetic_weights(L, seed=8, lamda=2.0)
    en = ProteinProductOfExperts.from_arrays(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    rng = np.random.default_rng(3)
    aa0 = np.tile(w.wt, (n, 1))
    for b in range(n):
        pos = rng.integers(0, L, size=b % 3)
        aa0[b, pos] = rng.integers(0, 20, size=pos.shape[0])
    x_host = port.aa_to_onehot(aa0)

    def run(pop, **kw):
        np.random.seed(0)
        args = argparse.Namespace(ppde_pas_length=2, nmut_threshold=0, paper_results=False, seed=6, ppde_verbose=False, **kw)
        return PPDE_PAS(args).run(pop, T, en, 0, L - 1, None, log_every=50)
    d = run(x_host.to("cuda:0"))
    h = run(x_host)
    r = run(torch.from_numpy(aa0), ppde_residue_io=True)
    assert d[0].device.type == "cuda" and h[0].device.type == "cpu" and r[0].device.type == "cpu"
    assert h[0].dtype == torch.float32 and tuple(h[0].shape) == (n, L, 20) and r[0].dtype == torch.uint8
    assert torch.equal(d[0].cpu(), h[0]) and np.array_equal(h[0].argmax(-1).numpy(), r[0].numpy())
    for k in (1, 2, 3, 4):
        assert np.array_equal(d[k], h[k]) and np.array_equal(d[k], r[k])


# ------------------------------------------------------------------------------------------- fused step kernels
def test_fused_step_equals_separate_kernels(monkeypatch):
    """Default step (Potts field update inside pas_propose, gradient combine inside pas_reverse_accept) against the same engine
    with the separate ppde_potts_incremental / cnn_grad_combine_sparse launches (PPDE_FUSE_POTTS=0, PPDE_FUSE_COMBINE=0): the
    fused kernels perform the same sums in the same order, so every row and every decision must agree BIT FOR BIT, across an
    exact-refresh iteration too."""
    from ppde_b200.engine import ChainEngine
    L, n, T, seed = 238, 96, 36, 11
    w = port.synthetic_weights(L, seed=2, lamda=15.0)
    m = _model(w)
    a = ChainEngine(m, n, 2, 0, False, seed=seed, num_steps=T)
    a.init_population(_wt_pop(m, w, n))
    monkeypatch.setenv("PPDE_FUSE_POTTS", "0")
    monkeypatch.setenv("PPDE_FUSE_COMBINE", "0")
    b = ChainEngine(m, n, 2, 0, False, seed=seed, num_steps=T)
    b.init_population(_wt_pop(m, w, n))
    assert a.fuse_potts and a.fuse_combine and not b.fuse_potts and not b.fuse_combine
    for t in range(T):
        a.run_steps(1, use_graph=True)
        b.run_steps(1, use_graph=(t % 2 == 0))          # graph replay and eager launches alternate on the reference side
        torch.cuda.synchronize()
        assert torch.equal(a.idx, b.idx) and torch.equal(a.accept, b.accept) and torch.equal(a.aa, b.aa), f"t={t}"
        assert torch.equal(a.row_cur, b.row_cur)
        rc, ry = a.row_cur.long(), a.rows_y.long()
        assert torch.equal(a.G[ry], b.G[ry]) and torch.equal(a.Gp[ry], b.Gp[ry]), f"t={t}: proposal rows differ"
        assert torch.equal(a.G[rc], b.G[rc]) and torch.equal(a.Gp[rc], b.Gp[rc])
        assert torch.equal(a.E_y, b.E_y) and torch.equal(a.lqr, b.lqr) and torch.equal(a.lqf, b.lqf)
    assert torch.equal(a.E_hist, b.E_hist) and torch.equal(a.best_aa, b.best_aa)


def test_long_sequence_kernels_vs_port():
    """L = 272 > 256: the position-per-thread PAS kernels, the fused Potts update / gradient combine and the tensor-core CNN do
    not apply; the strided shared-memory PAS kernels, ppde_potts_incremental and the fp32 SIMT CNN run instead.  A short run
    against the oracle port keeps that path covered."""
    from ppde_b200.engine import ChainEngine
    L, n, T, seed = 272, 6, 5, 3
    w = port.synthetic_weights(L, seed=4, lamda=3.0)
    m = _model(w)
    eng = ChainEngine(m, n, 2, 0, False, seed=seed, num_steps=T)
    eng.init_population(_wt_pop(m, w, n))
    assert not eng.fuse_potts and not eng.fuse_combine
    eng.run_steps(T, use_graph=True)
    torch.cuda.synchronize()
    en = port.PortEnergy(w)
    ref = port.PortSampler(2, 0, False, seed=seed).run(en.wt_onehot.repeat(n, 1, 1), T, en)
    assert _rel(eng.E_hist.cpu().numpy(), ref[3], abs(m.wt_H)) < 1e-4
    assert np.array_equal(eng.best_aa.cpu().numpy()[:, :L], ref[0].argmax(-1).numpy())
