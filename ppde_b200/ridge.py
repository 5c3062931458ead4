"""The oracle model of the reference (AugmentedLinearRegression, ppde/nets.py:315-347) on the B200 path.

    y(x) = mean_h [ W_h[0] * sqrt(1/reg_potts) * dH_potts(x) + sqrt(1/r_h) * <W_h[1:], onehot(x)> + b_h ]

The reference calls it at t = 0 and every `log_every` iterations on the whole population
(ppde/protein_samplers/ppde.py:48,156) and once on the final best samples (scripts/directed_evolution.py:83).
The 20 ridge heads are linear, so their mean is one head with averaged weights; the averaging is
done once on the host in float64 and the evaluation is a single kernel (`ppde_oracle_ridge`) that
reuses the Potts energy of the sampler's own field rows.
"""
import ctypes as C
import math
import os
import types

import numpy as np
import torch

from . import _lib
from . import weights as W
from .engine import PoEModel, Q, _ptr, _stream


class AugmentedLinearRegression:
    """Callable `x [n, L, 20] float one-hot -> [n]` with `.potts.index_list` as the reference's module has
    (scripts/directed_evolution.py:80-81 reads oracle.potts.index_list[0] / [-1])."""

    def __init__(self, model: PoEModel, coef, intercept, reg, reg_potts=1.0):
        coef = np.asarray(coef, dtype=np.float32).astype(np.float64)          # reference casts to float32 first (nets.py:327)
        icpt = np.asarray(intercept, dtype=np.float32).astype(np.float64).reshape(-1)
        reg = np.asarray(reg, dtype=np.float64).reshape(-1)
        if coef.ndim != 2 or coef.shape[1] != 1 + Q * model.L or len({coef.shape[0], icpt.shape[0], reg.shape[0]}) != 1:
            raise ValueError(f"ridge heads must be [H, 1 + 20*{model.L}]; got {coef.shape}")
        if not model.has_potts:
            raise ValueError("the oracle model needs the Potts expert (nets.py:318)")
        self.model = model
        self.lib = model.lib
        self.sbar = float(np.mean(coef[:, 0]) * math.sqrt(1.0 / float(reg_potts)))
        self.cbar = float(np.mean(icpt))
        wbar = np.mean(coef[:, 1:] * np.sqrt(1.0 / reg)[:, None], axis=0)
        self.wbar = torch.from_numpy(wbar.astype(np.float32)).to(model.device)
        lo = model.win_lo
        self.potts = types.SimpleNamespace(index_list=np.arange(lo, lo + model.Lp))

    @classmethod
    def from_dataset(cls, model, dataset_dir, reg_potts=1.0):
        coef, icpt, reg = W.load_ridge_heads(dataset_dir)
        return cls(model, coef, icpt, reg, reg_potts)

    def eval(self):
        return self

    def to(self, device):
        return self

    def score_states(self, aa, dH=None):
        """aa: uint8 device tensor [n, aa_stride]; dH: optional float tensor [n] with dH_potts of the same states."""
        m = self.model
        n = aa.shape[0]
        st = _stream()
        if dH is None:
            Gp = torch.empty(n, m.D, dtype=torch.float32, device=m.device)
            dH = torch.empty(n, dtype=torch.float32, device=m.device)
            m.potts_full(aa, n, _ptr(Gp), _ptr(dH), st)
        out = torch.empty(n, dtype=torch.float32, device=m.device)
        _lib.check(self.lib.ppde_oracle_ridge(_ptr(self.wbar), self.sbar, self.cbar, _ptr(aa), m.aa_stride, n, m.L,
                                              _ptr(dH), _ptr(out), st), "oracle_ridge")
        return out

    def score_engine(self, eng):
        """Current states of a ChainEngine, reusing the Potts field rows the sampler already holds."""
        m = self.model
        dH = torch.empty(eng.n, dtype=torch.float32, device=m.device)
        _lib.check(self.lib.ppde_potts_energy_rows(C.byref(m.potts), _ptr(eng.aa), m.aa_stride, eng.n, _ptr(eng.Gp), m.D,
                                                   _ptr(eng.row_cur), _ptr(dH), _stream()), "potts_energy_rows")
        return self.score_states(eng.aa, dH)

    def __call__(self, x):
        with torch.cuda.device(self.model.device):
            return self.score_states(self.model.onehot_to_aa(x))
