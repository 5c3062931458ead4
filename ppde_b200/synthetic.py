"""Synthetic model inputs for benchmarks and tests (SURVEY.md §8d "Synthetic inputs").

The reference's real potts.pkl files are not in its checkout (.MISSING_LARGE_BLOBS:3-5), and
protein lengths other than 96/104/237 have no shipped CNN checkpoint, so throughput runs use:
  Potts  J ~ N(0, 0.05^2) symmetrised (J[i,j,k,l] = J[j,i,l,k]), zero diagonal blocks, h ~ N(0, 0.5^2)
  CNN    OnehotCNN(20, 5, L) with torch's default Conv1d/Linear init bounds, seeds 0,1,2
  WT     uniform random residues (seeded)
Key layout matches what the reference reads (ppde/nets.py:247-251, 350-361).
"""
import math

import numpy as np


def synthetic_potts(Lp, seed=0, sigma_j=0.05, sigma_h=0.5, symmetric=True, zero_diag=True):
    rng = np.random.default_rng(seed)
    J = (rng.standard_normal((Lp, Lp, 20, 20)) * sigma_j).astype(np.float32)
    if symmetric:
        J = (0.5 * (J + J.transpose(1, 0, 3, 2))).astype(np.float32)
    if zero_diag:
        J[np.arange(Lp), np.arange(Lp)] = 0.0
    h = (rng.standard_normal((Lp, 20)) * sigma_h).astype(np.float32)
    return J, h


def synthetic_cnn(L, seeds=(0, 1, 2)):
    nets = []
    C = L
    b0 = 1 / math.sqrt(20 * 5); b1 = 1 / math.sqrt(C); b2 = 1 / math.sqrt(2 * C)
    for k in seeds:
        rng = np.random.default_rng(1000 + k)

        def uni(shape, bound):
            return ((rng.random(shape, dtype=np.float32) * 2 - 1) * bound).astype(np.float32)
        nets.append({"W0": uni((C, 20, 5), b0), "b0": uni((C,), b0), "W1": uni((2 * C, C), b1),
                     "b1": uni((2 * C,), b1), "d": uni((2 * C,), b2), "c": uni((1,), b2)})
    return nets


def synthetic_wt(L, seed=0):
    return np.random.default_rng(seed + 1000).integers(0, 20, size=L).astype(np.uint8)


def synthetic_problem(L, seed=0, window=None, sigma_j=0.05, sigma_h=0.5):
    """-> dict(wt, J, h, win_lo, cnn) for a length-L protein."""
    lo, hi = window if window is not None else (0, L - 1)
    J, h = synthetic_potts(hi - lo + 1, seed, sigma_j, sigma_h)
    return {"wt": synthetic_wt(L, seed), "J": J, "h": h, "win_lo": lo, "cnn": synthetic_cnn(L)}
