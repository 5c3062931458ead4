"""TEST INFRASTRUCTURE — CPU restatement ("port") of the reference's PPDE hot path.

This file is the ORACLE the CUDA path is checked against.  It is not product
code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  The product (ppde_b200/) never does and
fails loudly when its CUDA library is missing.

It restates, with PyTorch CPU ops in the same order as the reference so that the
arithmetic is bit-comparable, the following reference code:

  * Potts Hamiltonian and PoE energy + autograd gradient
      ppde/nets.py:282-299 (hamiltonian / forward(delta=True)),
      ppde/energy.py:97-108 (get_energy / get_energy_and_grads)
  * CNN ensemble expert   ppde/nets.py:363-376 (OnehotCNN.forward), :434-442 (mean)
  * edit distance / mask  ppde/utils.py:5-28
  * proposal probabilities ppde/utils.py:106-111 + torch.distributions.Categorical
      (__init__ renormalisation, sample -> multinomial, log_prob -> log(clamp(p)))
  * one MCMC iteration    ppde/protein_samplers/ppde.py:65-153
  * run()/6-tuple         ppde/protein_samplers/ppde.py:24-63,172-192
  * oracle model          ppde/nets.py:315-347 (AugmentedLinearRegression)

Parity pin: oracle/make_golden.py runs the UNMODIFIED reference (through
oracle/ref_harness.py, shared Philox streams) and stores its outputs under
tests/golden/; tests/test_oracle_golden.py checks this port against them.
Randomness is the indexed Philox stream of ppde_b200/philox.py.
"""
import math
import os
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F

from ppde_b200 import philox

ALPHABET = "ACDEFGHIKLMNPQRSTVWY"  # ppde/third_party/hsu/data_utils.py:48-72
EPS = torch.finfo(torch.float32).eps  # clamp_probs eps


def seq_to_aa(seq):
    return np.array([ALPHABET.index(c) for c in seq], dtype=np.uint8)


def aa_to_onehot(aa):
    """uint8 [n,L] -> float32 one-hot [n,L,20] (seqs_to_onehot, data_utils.py:150-157)."""
    aa = torch.as_tensor(np.asarray(aa).astype(np.int64))
    return F.one_hot(aa, 20).to(torch.float32)


def onehot_to_aa(x):
    return x.argmax(-1).to(torch.uint8).numpy()


@dataclass
class Weights:
    """Everything the hot path needs, as plain arrays (no reference objects)."""
    wt: np.ndarray                 # uint8 [L]
    J: np.ndarray                  # f32 [Lp,Lp,20,20]
    h: np.ndarray                  # f32 [Lp,20]
    win_lo: int                    # first Potts position (0-based, inclusive)
    cnn: list = field(default_factory=list)   # list of dicts: W0[C,20,5] b0[C] W1[2C,C] b1[2C] d[2C] c[1]
    lamda: float = 1.0
    reg_coef: float = 1.0
    ridge: list = field(default_factory=list)  # list of (coef f32 [1+20L], intercept f32, reg)

    @property
    def L(self):
        return int(self.wt.shape[0])

    @property
    def Lp(self):
        return int(self.J.shape[0])

    @property
    def win_hi(self):
        return self.win_lo + self.Lp - 1


class PortEnergy:
    """ProteinProductOfExperts restated (ppde/energy.py:72-108), potts branch."""

    def __init__(self, w: Weights):
        self.w = w
        self.J = torch.from_numpy(np.ascontiguousarray(w.J))
        self.h = torch.from_numpy(np.ascontiguousarray(w.h))
        self.lamda = w.lamda
        self.cnn = [{k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in net.items()}
                    for net in w.cnn]
        self.wt_onehot = aa_to_onehot(w.wt[None])          # [1,L,20]
        self.wt_H = self._hamiltonian(self._window(self.wt_onehot))

    def _window(self, x):                                   # nets.py:273-280
        return x[:, self.w.win_lo:self.w.win_hi + 1]

    def _hamiltonian(self, xw):                             # nets.py:282-290
        Jx = torch.einsum("ijkl,bjl->bik", self.J, xw)
        quad = torch.einsum("aik,aik->a", Jx, xw) / 2
        lin = (self.h[None] * xw).sum(-1).sum(-1)
        return quad + lin

    def potts_delta(self, x):                               # nets.py:292-297
        return self._hamiltonian(self._window(x)) - self.wt_H

    def _cnn_one(self, net, x):                             # nets.py:363-376
        z = F.relu(F.conv1d(x.transpose(1, 2), net["W0"], net["b0"]).transpose(1, 2))
        z = F.relu(F.linear(z, net["W1"], net["b1"]))
        z = torch.max(z, dim=1)[0]
        return F.linear(z, net["d"][None], net["c"])

    def fitness(self, x):                                   # nets.py:434-442
        preds = torch.stack([self._cnn_one(net, x) for net in self.cnn], 0)
        return torch.mean(preds, 0).squeeze()

    def get_energy(self, x):                                # energy.py:97-101
        fit = self.fitness(x)
        return self.potts_delta(x) + self.lamda * fit, fit

    def get_energy_and_grads(self, x):                      # energy.py:103-108
        x = x.detach().requires_grad_()
        fit = self.fitness(x)
        e = self.potts_delta(x) + self.lamda * fit
        g = torch.autograd.grad([e.sum()], x)[0]
        return e.detach(), fit.detach(), g


class PortOracleModel:
    """AugmentedLinearRegression restated (ppde/nets.py:315-347)."""

    def __init__(self, w: Weights, energy: PortEnergy):
        self.w, self.energy = w, energy
        self.coef = [torch.from_numpy(np.asarray(c, dtype=np.float32)) for c, _, _ in w.ridge]
        self.icpt = [torch.tensor([float(b)], dtype=torch.float32) for _, b, _ in w.ridge]
        self.reg = [float(r) for _, _, r in w.ridge]

    def __call__(self, x):
        dH = self.energy.potts_delta(x)
        flat = x.reshape(x.shape[0], -1)
        ys = []
        for W, b, r in zip(self.coef, self.icpt, self.reg):
            feat = torch.cat((math.sqrt(1 / self.w.reg_coef) * dH[..., None],
                              math.sqrt(1 / r) * flat), 1)
            ys.append((W * feat).sum(1) + b)
        return torch.stack(ys, 0).mean(0)


# ---------------------------------------------------------------- integer helpers
def edit_distance(aa, wt):
    """mut_distance on residue indices (ppde/utils.py:5-14). int64 [n]."""
    return (np.asarray(aa) != np.asarray(wt)[None]).sum(-1)


def revert_only_mask(aa, wt):
    """mutation_mask on residue indices (ppde/utils.py:17-28): bool [n,L,20],
    True = masked; only (i, wt_i) at positions with aa != wt stay False."""
    aa = np.asarray(aa); wt = np.asarray(wt)
    n, L = aa.shape
    mask = np.ones((n, L, 20), dtype=bool)
    b, i = np.nonzero(aa != wt[None])
    mask[b, i, wt[i]] = False
    return mask


def diversity_percent(aa):
    """unique sequences / K * 100 (scripts/make_figures.py:38-49)."""
    aa = np.asarray(aa)
    return len({bytes(r) for r in aa}) / aa.shape[0] * 100


def n_hops(aa, wt):
    """mean/std edit distance to WT (scripts/make_figures.py:29-36)."""
    d = edit_distance(aa, wt).astype(np.float64)
    return float(d.mean()), float(d.std())


# ---------------------------------------------------------------- proposal maths
def proposal_probs(logits):
    """safe_logits_to_probs (utils.py:106-111) followed by Categorical.__init__'s
    renormalisation. Returns p [n,20L] (masked entries end near eps/sum, not 0)."""
    z = logits - torch.logsumexp(logits, dim=-1, keepdim=True)
    p = torch.softmax(z, dim=-1).clamp(min=EPS, max=1 - EPS)
    return p / p.sum(-1, keepdim=True)


def log_prob_at(p, idx):
    """Categorical.log_prob: log(clamp_probs(p))[idx]."""
    return torch.log(p.clamp(min=EPS, max=1 - EPS)).gather(-1, idx[:, None])[:, 0]


def taylor_logits(grad, x):
    """(g - g[cur]) / 2 flattened to [n,20L] (ppde.py:98-100; ppde_temp = 2)."""
    s = grad - (grad * x).sum(-1).unsqueeze(-1)
    return s.reshape(x.shape[0], -1) / 2


@dataclass
class StepTrace:
    """Everything observable about one MCMC iteration (for parity tests)."""
    U: np.ndarray            # int32 [n]
    idx: np.ndarray          # int32 [S,n]   flat proposal index (i*20+a)
    lqf: np.ndarray          # f32 [S,n]
    lqr: np.ndarray          # f32 [S,n]
    margin: np.ndarray       # f32 [S,n]   1 - second/best of the exponential race
    aa_y: np.ndarray         # uint8 [n,L]
    e_x: np.ndarray; fit_x: np.ndarray
    e_y: np.ndarray; fit_y: np.ndarray
    log_acc: np.ndarray      # f32 [n]
    u_acc: np.ndarray        # f32 [n]
    accept: np.ndarray       # bool [n]
    aa_rec: np.ndarray       # uint8 [n,L] state recorded in history (post-accept, pre-reset)
    aa_new: np.ndarray       # uint8 [n,L] state carried to t+1 (post-reset)
    e_new: np.ndarray; fit_new: np.ndarray
    grad_x: np.ndarray = None
    grad_y: np.ndarray = None


class PortSampler:
    """PPDE_PAS restated (ppde/protein_samplers/ppde.py:8-192) on indexed streams.

    `fixed_S=False` reproduces the reference's data-dependent `max_u` loop bound
    (ppde.py:68,83); `fixed_S=True` always runs S_max = 2*pas-1 sub-steps, which is
    what the CUDA path does — equivalent because sub-steps s >= U contribute nothing
    and streams are indexed, not consumed.
    """

    def __init__(self, pas_length=2, nmut_threshold=0, paper_results=False, seed=0,
                 chain_offset=0, fixed_S=True):
        self.pas = int(pas_length)
        self.thr = int(nmut_threshold) if nmut_threshold else np.iinfo(np.int32).max
        self.paper = bool(paper_results)
        self.seed = seed
        self.chain_offset = chain_offset
        self.fixed_S = fixed_S

    def step(self, energy: PortEnergy, t, cur_x, x_anchor, keep_grads=False):
        """One iteration t. cur_x: one-hot [n,L,20]; x_anchor: the paper-mode `x`
        (initial population) — ignored in hard mode. Returns StepTrace."""
        w = energy.w
        n, L = cur_x.shape[0], cur_x.shape[1]
        chains = np.arange(n, dtype=np.uint32) + np.uint32(self.chain_offset)
        wt = w.wt
        U = philox.path_lengths(self.seed, t, chains, self.pas)           # ppde.py:67
        S = 2 * self.pas - 1 if self.fixed_S else int(U.max())
        u_mask = torch.from_numpy((np.arange(S)[None, :] < U[:, None]).astype(np.float32))
        x = x_anchor if self.paper else cur_x.clone()                      # :76-77
        e_x, fit_x, g_x = energy.get_energy_and_grads(cur_x)               # :79
        pos_mask = torch.ones(n, L, 20, dtype=torch.bool)
        pos_mask[:, w.win_lo:w.win_hi + 1] = False                         # :59-63 (min_pos,max_pos)
        pos_mask = pos_mask.reshape(n, -1)

        idxs, lqf, margins, traj, fwd_p = [], [], [], [], []
        cur = cur_x.detach()
        for s in range(S):
            aa = onehot_to_aa(cur)
            at_thr = torch.from_numpy(edit_distance(aa, wt) >= self.thr)   # :86-91
            mask = torch.from_numpy(revert_only_mask(aa, wt)).reshape(n, -1)
            mask[~at_thr] = False                                          # :95
            logits = taylor_logits(g_x, cur)                               # :98-100
            logits[mask] = -np.inf                                         # :103
            logits[pos_mask] = -np.inf                                     # :104
            p = proposal_probs(logits)                                     # :106-107
            u = torch.from_numpy(philox.proposal_uniforms(self.seed, t, s, chains, 20 * L))
            race = p / (-torch.log(u))                                     # multinomial
            top2 = torch.topk(race, 2, dim=-1).values
            idx = torch.argmax(race, dim=-1)
            margins.append((1 - top2[:, 1] / top2[:, 0]).numpy())
            idxs.append(idx); fwd_p.append(p); traj.append(cur)
            lqf.append(log_prob_at(p, idx))
            move = F.one_hot(idx, 20 * L).to(torch.float32).reshape(n, L, 20)  # :111-115
            row = move.sum(-1, keepdim=True)
            new = cur * (1.0 - row) + move
            m = u_mask[:, s].reshape(n, 1, 1)
            cur = m * new + (1 - m) * cur
        y = cur
        e_y, fit_y, g_y = energy.get_energy_and_grads(y)                   # :118-120
        traj.append(y)
        log_ratio = torch.zeros(n)
        lqr = []
        for s in range(S):                                                 # :124-132
            q = proposal_probs(taylor_logits(g_y, traj[s + 1]))
            lqr.append(log_prob_at(q, idxs[s]))
            log_ratio = log_ratio + u_mask[:, s] * (lqr[-1] - lqf[s])
        log_acc = (e_y - e_x) + log_ratio                                  # :135-136
        u_acc = torch.from_numpy(philox.accept_uniforms(self.seed, t, chains))
        acc = (log_acc.exp() >= u_acc).float()                             # :138
        a3 = acc.reshape(n, 1, 1)
        rec = y * a3 + (1.0 - a3) * x                                      # :139
        e_new = e_y * acc + e_x * (1.0 - acc)                              # :141
        fit_new = fit_y * acc + fit_x * (1.0 - acc)                        # :143
        nxt = rec.clone()
        if not self.paper:                                                 # :148-153
            over = torch.from_numpy(edit_distance(onehot_to_aa(nxt), wt) >= self.thr)
            nxt[over] = energy.wt_onehot[0]
        return StepTrace(
            U=U, idx=torch.stack(idxs).numpy().astype(np.int32),
            lqf=torch.stack(lqf).numpy(), lqr=torch.stack(lqr).numpy(),
            margin=np.stack(margins), aa_y=onehot_to_aa(y),
            e_x=e_x.numpy(), fit_x=fit_x.numpy(), e_y=e_y.numpy(), fit_y=fit_y.numpy(),
            log_acc=log_acc.numpy(), u_acc=u_acc.numpy(), accept=acc.numpy() > 0.5,
            aa_rec=onehot_to_aa(rec), aa_new=onehot_to_aa(nxt),
            e_new=e_new.numpy(), fit_new=fit_new.numpy(),
            grad_x=g_x.numpy() if keep_grads else None,
            grad_y=g_y.numpy() if keep_grads else None), nxt

    def run(self, initial_population, num_steps, energy: PortEnergy, random_idx=0, traces=None,
            cpu_alias_quirk=False):
        """Whole run; returns the reference's 6-tuple (ppde.py:172-192).

        `cpu_alias_quirk`: on a CPU device the reference's `cur_x.detach().cpu().numpy()`
        (ppde.py:142,146) is a VIEW of `cur_x`, so the in-place hard reset at :153 also
        rewrites the history entry just appended (best_x / random_traj then hold the
        post-reset WT state while best_energy is the pre-reset energy).  On a CUDA device
        `.cpu()` copies and the history keeps the pre-reset state.  The goldens were
        produced on CPU, so the pin test sets this flag; the CUDA path follows the CUDA
        behaviour (flag off)."""
        x0 = initial_population.clone()
        with torch.no_grad():
            e0, f0 = energy.get_energy(x0)                                 # :41-42
        e_hist, f_hist = [e0], [f0]
        all_aa = [onehot_to_aa(x0)]
        cur = x0.clone()
        for t in range(num_steps):
            tr, cur = self.step(energy, t, cur, x0)
            e_hist.append(torch.from_numpy(tr.e_new)); f_hist.append(torch.from_numpy(tr.fit_new))
            all_aa.append(tr.aa_new if cpu_alias_quirk else tr.aa_rec)
            if traces is not None:
                traces.append(tr)
        e_hist = torch.stack(e_hist); f_hist = torch.stack(f_hist)
        best_e, best_t = torch.max(e_hist, 0)                              # :173 first occurrence
        all_aa = np.stack(all_aa, 0)
        n = x0.shape[0]
        best_aa = all_aa[best_t.numpy(), np.arange(n)]
        best_f = f_hist[best_t, torch.arange(n)]
        traj = [aa_to_onehot(a[random_idx][None])[0].numpy() for a in all_aa]
        return (aa_to_onehot(best_aa), best_e.numpy(), best_f.numpy(),
                e_hist.numpy(), f_hist.numpy(), traj)


from ppde_b200.synthetic import synthetic_potts, synthetic_cnn, synthetic_problem  # noqa: E402,F401


def synthetic_weights(L, seed=0, lamda=1.0, window=None, sigma_j=0.05, sigma_h=0.5):
    """Weights for the SURVEY.md §8d synthetic problem (same arrays the CUDA path is given)."""
    pr = synthetic_problem(L, seed, window, sigma_j, sigma_h)
    return Weights(wt=pr["wt"], J=pr["J"], h=pr["h"], win_lo=pr["win_lo"], cnn=pr["cnn"], lamda=lamda)


GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def golden_weights(prot, window, lamda, potts_seed=0, sigma_j=0.05, sigma_h=0.5,
                   symmetric=True, zero_diag=True):
    """Weights from the committed fixtures (tests/golden/weights_<PROT>.npz: shipped WT,
    CNN checkpoints, ridge heads) + regenerated synthetic Potts parameters."""
    import os
    z = np.load(os.path.join(GOLDEN_DIR, f"weights_{prot}.npz"))
    wt = seq_to_aa(str(z["wt_seq"]))
    lo, hi = int(window[0]), int(window[1])
    J, h = synthetic_potts(hi - lo + 1, int(potts_seed), float(sigma_j), float(sigma_h),
                           bool(symmetric), bool(zero_diag))
    cnn = [{k: z[f"cnn{i}_{k}"] for k in ("W0", "b0", "W1", "b1", "d", "c")} for i in range(3)]
    ridge = [(z["ridge_coef"][i], z["ridge_intercept"][i], z["ridge_reg"][i]) for i in range(20)]
    return Weights(wt=wt, J=J, h=h, win_lo=lo, cnn=cnn, lamda=float(lamda), ridge=ridge)
