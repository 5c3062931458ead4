"""TEST INFRASTRUCTURE — generates tests/golden/* by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Outputs (committed):
  tests/golden/weights_<PROT>.npz   shipped fixtures repacked as plain arrays: WT residues,
                                    the 3 CNN checkpoints, the 20 ridge heads (f32, as the
                                    reference casts them: ppde/nets.py:327-328)
  tests/golden/kat_int.npz          mut_distance / mutation_mask / safe_logits_to_probs KATs
  tests/golden/kat_energy_<case>.npz  get_energy_and_grads on random mutants
  tests/golden/traj_<case>.npz      full PPDE_PAS.run on shared Philox streams, with
                                    per-iteration internals captured by spies (no reference
                                    code is modified; spies wrap the energy object and the
                                    torch RNG entry points only)
Synthetic Potts parameters are regenerated from (seed, window, sigmas) stored in each file;
a float64 checksum of J guards against generator drift.
"""
import contextlib
import io
import os
import pickle
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)
from oracle import ref_harness as rh  # noqa: E402

GOLD = os.path.join(REPO, "tests", "golden")

TRAJ_CASES = {
    # name: protein, window, n, lamda, pas, nmut, paper, T, seed, potts kwargs
    "pabp_hard": dict(prot="PABP", window=(3, 92), n=8, lamda=5.0, pas=2, nmut=0, paper=False, T=12, seed=11),
    "pabp_thr3": dict(prot="PABP", window=(3, 92), n=8, lamda=5.0, pas=2, nmut=3, paper=False, T=16, seed=12),
    "pabp_paper_thr3_pas3": dict(prot="PABP", window=(0, 95), n=6, lamda=5.0, pas=3, nmut=3, paper=True, T=10, seed=13),
    "ube4b_cfg2": dict(prot="UBE4B", window=(22, 97), n=8, lamda=0.5, pas=2, nmut=10, paper=False, T=10, seed=14),
    "gfp_cfg3": dict(prot="GFP", window=(0, 236), n=4, lamda=15.0, pas=2, nmut=0, paper=False, T=4, seed=15),
    "gfp_cfg4_paper_pas10": dict(prot="GFP", window=(2, 230), n=3, lamda=15.0, pas=10, nmut=10, paper=True, T=3, seed=16),
}

ENERGY_CASES = {
    "pabp": dict(prot="PABP", window=(3, 92), lamda=5.0, potts=dict()),
    "pabp_asym": dict(prot="PABP", window=(5, 80), lamda=5.0, potts=dict(symmetric=False, zero_diag=False)),
    "ube4b": dict(prot="UBE4B", window=(22, 97), lamda=0.5, potts=dict()),
    "gfp": dict(prot="GFP", window=(0, 236), lamda=15.0, potts=dict()),
}


def repack_weights(key):
    protein = rh.PROTEINS[key]
    src = os.path.join(rh.REFERENCE_ROOT, "weights", protein)
    with open(os.path.join(src, "wt.fasta")) as fh:
        lines = fh.read().split("\n")
    fasta_id = lines[0][1:].split()[0]
    seq = "".join(l.strip() for l in lines[1:])
    out = {"fasta_id": np.array(fasta_id), "wt_seq": np.array(seq)}
    for k in range(3):
        sd = torch.load(os.path.join(src, f"onehot_cnn_seed={k}.pt"), map_location="cpu")["model"]
        out[f"cnn{k}_W0"] = sd["encoder.weight"].numpy()
        out[f"cnn{k}_b0"] = sd["encoder.bias"].numpy()
        out[f"cnn{k}_W1"] = sd["embedding.0.weight"].numpy()
        out[f"cnn{k}_b1"] = sd["embedding.0.bias"].numpy()
        out[f"cnn{k}_d"] = sd["decoder.weight"].numpy()[0]
        out[f"cnn{k}_c"] = sd["decoder.bias"].numpy()
    coefs, icpts, regs = [], [], []
    for seed in range(20):
        with open(os.path.join(src, f"results-predictor=ev+onehot-train=-1-seed={seed}-linear.pkl"), "rb") as fh:
            r = pickle.load(fh)
        coefs.append(torch.from_numpy(r["coef_"]).float().numpy())
        icpts.append(torch.FloatTensor([r["intercept_"]]).numpy()[0])
        regs.append(float(r["reg_coef"]))
    out["ridge_coef"] = np.stack(coefs)
    out["ridge_intercept"] = np.array(icpts, dtype=np.float32)
    out["ridge_reg"] = np.array(regs, dtype=np.float64)
    np.savez_compressed(os.path.join(GOLD, f"weights_{key}.npz"), **out)


def kat_int():
    rh.install_shims()
    from ppde.utils import mut_distance, mutation_mask, safe_logits_to_probs
    rng = np.random.default_rng(5)
    L, n = 31, 9
    wt = rng.integers(0, 20, L)
    aa = np.tile(wt, (n, 1))
    for b in range(1, n):                       # chain 0 = WT; chain b has <= 2b edits
        pos = rng.integers(0, L, 2 * b)
        aa[b, pos] = rng.integers(0, 20, 2 * b)
    oh = torch.nn.functional.one_hot(torch.from_numpy(aa), 20).float()
    wt_oh = torch.nn.functional.one_hot(torch.from_numpy(wt), 20).float()[None]
    dist = mut_distance(oh, wt_oh).numpy()
    mask = mutation_mask(oh, wt_oh).numpy()
    logits = torch.from_numpy((rng.standard_normal((n, L * 20)) * 6).astype(np.float32))
    logits[2, 40:300] = -np.inf
    logits[3, :] = -np.inf
    logits[3, 17] = 0.0
    logits[4] *= 8.0                             # very peaked: many entries under eps
    p = safe_logits_to_probs(logits)
    cat = torch.distributions.one_hot_categorical.OneHotCategorical(probs=p)
    pn = cat._categorical.probs
    idx = torch.from_numpy(rng.integers(0, L * 20, n))
    lp = cat.log_prob(torch.nn.functional.one_hot(idx, L * 20).float())
    np.savez_compressed(os.path.join(GOLD, "kat_int.npz"), wt=wt.astype(np.uint8), aa=aa.astype(np.uint8),
                        dist=dist, mask=mask, logits=logits.numpy(), probs_safe=p.numpy(),
                        probs_norm=pn.numpy(), idx=idx.numpy(), log_prob=lp.numpy())


def _potts_meta(case, seed=0):
    kw = dict(symmetric=True, zero_diag=True, sigma_j=0.05, sigma_h=0.5)
    kw.update(case.get("potts", {}))
    return dict(potts_seed=seed, **kw)


def kat_energy(name, case, tmp):
    meta = _potts_meta(case)
    root, protein = rh.make_weights_dir(os.path.join(tmp, "e_" + name), case["prot"], window=case["window"],
                                        seed=meta["potts_seed"], symmetric=meta["symmetric"],
                                        zero_diag=meta["zero_diag"], sigma_j=meta["sigma_j"], sigma_h=meta["sigma_h"])
    n = 6
    args, energy, sampler, oracle = rh.load_reference(root, protein, n, case["lamda"])
    wt = energy.wt_onehot.argmax(-1)[0].numpy()
    L = wt.shape[0]
    rng = np.random.default_rng(99)
    aa = np.tile(wt, (n, 1))
    for b in range(1, n):
        k = [0, 1, 2, 5, 12, 30][b]
        pos = rng.choice(L, k, replace=False)
        aa[b, pos] = rng.integers(0, 20, k)
    aa[5, 10:16] = aa[5, 40:46]                 # repeated 5-mers -> exact max-pool ties
    x = torch.nn.functional.one_hot(torch.from_numpy(aa), 20).float().requires_grad_()
    e, fit, g = energy.get_energy_and_grads(x)
    with torch.no_grad():
        e2, fit2 = energy.get_energy(x.detach())
        orc = oracle(x.detach())
        pot = energy.unsupervised_expert(energy.unsupervised_expert.preprocess_onehot(x.detach()), delta=True)
    J = energy.unsupervised_expert.J.detach().numpy()
    np.savez_compressed(
        os.path.join(GOLD, f"kat_energy_{name}.npz"), prot=np.array(case["prot"]),
        window=np.array(case["window"]), lamda=np.float64(case["lamda"]),
        aa=aa.astype(np.uint8), e=e.detach().numpy(), fit=fit.detach().numpy(), grad=g.numpy(),
        e_nograd=e2.numpy(), fit_nograd=fit2.numpy(), oracle=orc.numpy(), potts_delta=pot.numpy(),
        wt_H=energy.unsupervised_expert.wt_H.detach().numpy(),
        J_checksum=np.float64(J.astype(np.float64).sum()), J_abs_checksum=np.float64(np.abs(J).astype(np.float64).sum()),
        **{k: np.array(v) for k, v in meta.items()})


def traj(name, case, tmp):
    meta = _potts_meta(case)
    root, protein = rh.make_weights_dir(os.path.join(tmp, "t_" + name), case["prot"], window=case["window"],
                                        seed=meta["potts_seed"])
    n, T = case["n"], case["T"]
    args, energy, sampler, oracle = rh.load_reference(root, protein, n, case["lamda"], case["pas"],
                                                      case["nmut"], case["paper"])
    spy = rh.EnergySpy(energy)
    rec = {}
    lp_log = []
    OHC = torch.distributions.one_hot_categorical.OneHotCategorical
    orig_lp = OHC.log_prob

    def spy_lp(self, value):
        out = orig_lp(self, value)
        lp_log.append(out.detach().numpy().copy())
        return out

    pop = energy.wt_onehot.repeat(n, 1, 1)
    np.random.seed(case["seed"])                 # fixes random_idx (ppde.py:37)
    OHC.log_prob = spy_lp
    try:
        with rh.SharedStreams(case["seed"], n, record=rec), contextlib.redirect_stdout(io.StringIO()):
            out = sampler.run(pop, T, spy, int(oracle.potts.index_list[0]),
                              int(oracle.potts.index_list[-1]), oracle, 5)
    finally:
        OHC.log_prob = orig_lp
    np.random.seed(case["seed"])
    random_idx = np.random.randint(0, n)
    best_x, best_e, best_f, e_hist, f_hist, rtraj = out
    S_max = 2 * case["pas"] - 1
    idx = np.full((T, S_max, n), -1, dtype=np.int32)
    lqf = np.full((T, S_max, n), np.nan, dtype=np.float32)
    lqr = np.full((T, S_max, n), np.nan, dtype=np.float32)
    k = 0
    for t in range(T):
        mu = len(rec["idx"][t])
        for s in range(mu):
            idx[t, s] = rec["idx"][t][s]
            lqr[t, s] = lp_log[k]; lqf[t, s] = lp_log[k + 1]   # ppde.py:132 evaluates reverse first
            k += 2
    assert k == len(lp_log)
    aa_x = np.stack([c["aa"] for c in spy.calls[0::2]])       # state at the start of iteration t
    aa_y = np.stack([c["aa"] for c in spy.calls[1::2]])
    J = energy.unsupervised_expert.J.detach().numpy()
    np.savez_compressed(
        os.path.join(GOLD, f"traj_{name}.npz"), prot=np.array(case["prot"]), window=np.array(case["window"]),
        n=n, T=T, lamda=np.float64(case["lamda"]), pas=case["pas"], nmut=case["nmut"], paper=case["paper"],
        seed=case["seed"], random_idx=random_idx,
        U=np.stack(rec["U"]), idx=idx, lqf=lqf, lqr=lqr, u_acc=np.stack(rec["u_acc"]),
        aa_x=aa_x, aa_y=aa_y,
        e_x=np.stack([c["e"] for c in spy.calls[0::2]]), fit_x=np.stack([c["fit"] for c in spy.calls[0::2]]),
        e_y=np.stack([c["e"] for c in spy.calls[1::2]]), fit_y=np.stack([c["fit"] for c in spy.calls[1::2]]),
        best_aa=best_x.argmax(-1).numpy().astype(np.uint8), best_e=best_e, best_f=best_f,
        e_hist=e_hist, f_hist=f_hist,
        random_traj_aa=np.stack([r.argmax(-1) for r in rtraj]).astype(np.uint8),
        J_checksum=np.float64(J.astype(np.float64).sum()),
        **{kk: np.array(v) for kk, v in meta.items()})


def main():
    assert rh.reference_available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="ppde_golden_")
    try:
        for key in rh.PROTEINS:
            repack_weights(key)
        kat_int()
        for name, case in ENERGY_CASES.items():
            kat_energy(name, case, tmp)
            print("energy KAT", name)
        for name, case in TRAJ_CASES.items():
            traj(name, case, tmp)
            print("trajectory", name)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
