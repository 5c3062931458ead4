"""Drop-in energy objects with the reference's interface (ppde/energy.py:71-164).

    ProteinProductOfExperts(args)   E = dH_potts(x) + lamda * mean_k CNN_k(x)
    ProteinSupervised(args)         E = mean_k CNN_k(x)

Both expose what the samplers use: `.get_energy(x) -> (e, fit)`,
`.get_energy_and_grads(x) -> (e, fit, grad_x)`, `.wt_onehot [1,L,20]`, `.lamda`, `.to(device)`.
`x` is the reference's float one-hot [n, L, 20]; all arithmetic runs in the CUDA kernels.
"""
import os

import numpy as np
import torch

from . import weights as W
from .engine import PoEModel, Q


class _EnergyBase:
    model: PoEModel

    def _finish(self):
        m = self.model
        wt = torch.from_numpy(m.wt_host.astype(np.int64))
        self.wt_onehot = torch.nn.functional.one_hot(wt, Q).float()[None].to(m.device)     # energy.py:95
        self.wt_aa = m.wt_host

    def to(self, device):                       # reference call site: energy_func.to(args.device)
        if torch.device(device).type != "cuda":
            raise RuntimeError("ppde_b200 energies live on a CUDA device; there is no CPU fallback")
        return self

    def eval(self):
        return self

    def get_energy(self, x):                    # energy.py:97-101 / :153-155
        aa = self.model.onehot_to_aa(x)
        e, fit, _, _ = self.model.energy(aa, want_grad=False)
        return (e, fit) if self.model.has_potts else (fit, fit)

    def get_energy_and_grads(self, x):          # energy.py:103-108 / :157-160
        aa = self.model.onehot_to_aa(x)
        e, fit, g, _ = self.model.energy(aa, want_grad=True)
        return (e, fit, g) if self.model.has_potts else (fit, fit, g)

    def get_supervised_expert(self, x):         # energy.py:134-136
        return self.get_energy(x)[1]


class ProteinProductOfExperts(_EnergyBase):
    """Reads the same files and `args` attributes as the reference constructor (energy.py:72-95):
    args.energy_lamda, args.unsupervised_expert ('potts'), args.protein_weights, args.protein, args.device."""

    def __init__(self, args):
        if getattr(args, "unsupervised_expert", "potts") != "potts":
            raise NotImplementedError("only the Potts unsupervised expert is on the B200 hot path "
                                      "(transformer experts are out of scope, SURVEY.md §8)")
        dataset = os.path.join(args.protein_weights, args.protein)
        seqs, ids = W.read_fasta(os.path.join(dataset, "wt.fasta"))
        potts = W.load_potts(dataset, ids[0])
        cnn = W.load_cnn_ensemble(dataset)
        self.model = PoEModel(W.seq_to_aa(seqs[0]), potts["J"], potts["h"], potts["win_lo"], cnn,
                              args.energy_lamda, device=getattr(args, "device", None))
        self.lamda = args.energy_lamda
        self.reg_coef = potts["reg_coef"]
        self._finish()

    @classmethod
    def from_arrays(cls, wt_aa, J, h, win_lo, cnn, lamda, device=None):
        self = cls.__new__(cls)
        self.model = PoEModel(wt_aa, J, h, win_lo, cnn, lamda, device=device)
        self.lamda = lamda
        self.reg_coef = 1.0
        self._finish()
        return self

    @classmethod
    def from_reference(cls, ref, device=None):
        """Ingest a constructed reference `ppde.energy.ProteinProductOfExperts` (weights are read from
        its modules: .unsupervised_expert.{J,bias,index_list}, .supervised_expert.surrogates, .wt_onehot)."""
        potts = ref.unsupervised_expert
        wt_aa = ref.wt_onehot[0].argmax(-1).cpu().numpy().astype(np.uint8)
        cnn = [W.cnn_from_state_dict(s.state_dict()) for s in ref.supervised_expert.surrogates]
        self = cls.from_arrays(wt_aa, potts.J.detach().cpu().numpy(), potts.bias.detach().cpu().numpy(),
                               int(potts.index_list[0]), cnn, float(ref.lamda), device=device)
        self.reg_coef = float(getattr(potts, "reg_coef", 1.0))
        return self

    def get_unsupervised_expert(self, x):       # energy.py:138-140
        aa = self.model.onehot_to_aa(x)
        return self.model.energy(aa, want_grad=False)[3]


class ProteinSupervised(_EnergyBase):
    """CNN-ensemble-only energy (ppde/energy.py:143-164)."""

    def __init__(self, args):
        dataset = os.path.join(args.protein_weights, args.protein)
        seqs, _ = W.read_fasta(os.path.join(dataset, "wt.fasta"))
        self.model = PoEModel(W.seq_to_aa(seqs[0]), None, None, 0, W.load_cnn_ensemble(dataset), 1.0,
                              device=getattr(args, "device", None))
        self.lamda = 1.0
        self._finish()

    @classmethod
    def from_arrays(cls, wt_aa, cnn, device=None):
        self = cls.__new__(cls)
        self.model = PoEModel(wt_aa, None, None, 0, cnn, 1.0, device=device)
        self.lamda = 1.0
        self._finish()
        return self


def as_b200_energy(energy_function, device=None):
    """Accept either one of this module's energies or a reference energy object."""
    if isinstance(energy_function, _EnergyBase):
        return energy_function
    if hasattr(energy_function, "unsupervised_expert") and hasattr(energy_function, "supervised_expert"):
        return ProteinProductOfExperts.from_reference(energy_function, device=device)
    raise TypeError("energy_function must be a ppde_b200 energy or a reference ProteinProductOfExperts")
