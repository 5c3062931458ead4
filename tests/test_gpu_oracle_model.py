"""Oracle model (AugmentedLinearRegression, ppde/nets.py:315-347) on the CUDA path vs the port on the shipped
ridge heads (tests/golden/weights_*.npz), and through the sampler's log_every fast path."""
import numpy as np
import pytest
import torch

from oracle import ppde_port as port

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prot,window", [("PABP", (3, 90)), ("UBE4B", (22, 97)), ("GFP", (0, 236))])
def test_oracle_model_matches_port(prot, window):
    from ppde_b200.engine import ChainEngine, PoEModel
    from ppde_b200.ridge import AugmentedLinearRegression

    w = port.golden_weights(prot, window, lamda=1.0)
    m = PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    coef = np.stack([c for c, _, _ in w.ridge]); icpt = np.array([b for _, b, _ in w.ridge]); reg = np.array([r for _, _, r in w.ridge])
    om = AugmentedLinearRegression(m, coef, icpt, reg, reg_potts=w.reg_coef)
    assert om.potts.index_list[0] == w.win_lo and om.potts.index_list[-1] == w.win_lo + w.J.shape[0] - 1
    rng = np.random.default_rng(7)
    n, L = 48, len(w.wt)
    aa = np.tile(w.wt, (n, 1)).astype(np.uint8)
    for b in range(1, n):
        pos = rng.integers(0, L, size=b % 12 + 1)
        aa[b, pos] = rng.integers(0, 20, size=pos.shape[0])
    x = port.aa_to_onehot(aa)
    ref = port.PortOracleModel(w, port.PortEnergy(w))(x).numpy()
    got = om(x.to(m.device)).cpu().numpy()
    scale = max(np.abs(ref).max(), 1e-6)
    assert np.abs(got - ref).max() <= 1e-4 * scale, np.abs(got - ref).max() / scale
    # fast path: scores straight from a sampler's field rows (after a few iterations: mixed own / WT rows)
    eng = ChainEngine(m, n, pas_length=2, nmut_threshold=5, seed=3, num_steps=4)
    pad = np.zeros((n, m.aa_stride), dtype=np.uint8); pad[:, :L] = aa
    eng.init_population(torch.from_numpy(pad).to(m.device))
    eng.run_steps(4, use_graph=False)
    torch.cuda.synchronize()
    cur = eng.aa.cpu().numpy()[:, :L]
    ref2 = port.PortOracleModel(w, port.PortEnergy(w))(port.aa_to_onehot(cur)).numpy()
    got2 = om.score_engine(eng).cpu().numpy()
    # the incrementally updated field carries ~1e-7 drift per update; the oracle's dH is small next to |H(wt)|
    tol = 1e-4 * max(np.abs(ref2).max(), abs(om.sbar) * abs(m.wt_H), 1e-6)
    assert np.abs(got2 - ref2).max() <= tol, (np.abs(got2 - ref2).max(), tol)


@pytest.mark.parametrize("name", ["pabp", "pabp_asym", "ube4b", "gfp"])
def test_oracle_model_vs_reference_kat(name):
    """Against the UNMODIFIED reference's own oracle outputs (tests/golden/kat_energy_*.npz, key 'oracle')."""
    import os
    from ppde_b200.engine import PoEModel
    from ppde_b200.ridge import AugmentedLinearRegression

    z = np.load(os.path.join(port.GOLDEN_DIR, f"kat_energy_{name}.npz"))
    meta = dict(potts_seed=int(z["potts_seed"]), sigma_j=float(z["sigma_j"]), sigma_h=float(z["sigma_h"]),
                symmetric=bool(z["symmetric"]), zero_diag=bool(z["zero_diag"]))
    w = port.golden_weights(str(z["prot"]), z["window"], float(z["lamda"]), **meta)
    m = PoEModel(w.wt, w.J, w.h, w.win_lo, w.cnn, w.lamda, device="cuda:0")
    coef = np.stack([c for c, _, _ in w.ridge]); icpt = np.array([b for _, b, _ in w.ridge]); reg = np.array([r for _, _, r in w.ridge])
    om = AugmentedLinearRegression(m, coef, icpt, reg, reg_potts=w.reg_coef)
    aa = z["aa"]
    pad = np.zeros((aa.shape[0], m.aa_stride), dtype=np.uint8); pad[:, :aa.shape[1]] = aa
    got = om.score_states(torch.from_numpy(pad).to(m.device)).cpu().numpy()
    ref = np.asarray(z["oracle"], dtype=np.float64)
    scale = max(np.abs(ref).max(), abs(om.sbar) * abs(float(z["wt_H"])), 1e-6)
    assert np.abs(got - ref).max() <= 1e-4 * scale, np.abs(got - ref).max() / scale
