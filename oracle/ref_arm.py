"""TEST INFRASTRUCTURE — the reference arm: times the UNMODIFIED reference implementation of the hot path
(`ppde.protein_samplers.ppde.PPDE_PAS.run` + `ppde.energy.ProteinProductOfExperts`, ppde/protein_samplers/ppde.py:24-192,
ppde/energy.py:72-108) on the host cores, on the same synthetic problem the CUDA arm runs.

The package is imported from oracle/_ref (staged by oracle/stage_ref.py; it travels to the GPU box) under the two module
shims of SURVEY.md §8c (`Bio.SeqIO`, `esm_one_hot`: un-installed imports the path never calls).  The synthetic weights are
written in the reference's own on-disk formats (wt.fasta, potts.pkl, onehot_cnn_seed=k.pt: ppde/nets.py:247-262, 416-424,
ppde/energy.py:91-95) into a temporary directory and read by the reference's own constructors.

Only bench.py (`--impl reference`, `cpu_baseline`) and tests may import this module.
"""
import argparse
import contextlib
import io
import os
import pickle
import sys
import tempfile
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
ALPHABET = "ACDEFGHIKLMNPQRSTVWY"
PROTEIN = "SYNTHETIC"


def available():
    return os.path.isfile(os.path.join(REF_DIR, "ppde", "protein_samplers", "ppde.py"))


def _import_reference():
    from oracle import ref_harness as rh
    rh.REFERENCE_ROOT = REF_DIR
    rh.install_shims()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    from ppde.energy import ProteinProductOfExperts
    from ppde.protein_samplers.ppde import PPDE_PAS
    return ProteinProductOfExperts, PPDE_PAS


def write_weights_dir(root, pr):
    """pr: ppde_b200.synthetic.synthetic_problem dict (wt, J, h, win_lo, cnn) -> <root>/SYNTHETIC/{wt.fasta,potts.pkl,*.pt}"""
    d = os.path.join(root, PROTEIN)
    os.makedirs(d, exist_ok=True)
    L = int(pr["wt"].shape[0])
    with open(os.path.join(d, "wt.fasta"), "w") as fh:
        fh.write(">SYNTHETIC\n" + "".join(ALPHABET[a] for a in pr["wt"]) + "\n")
    Lp = int(pr["J"].shape[0])
    index_list = np.arange(pr["win_lo"], pr["win_lo"] + Lp, dtype=np.int64) + 1        # fasta id without '/': offset 1 (nets.py:257-261)
    with open(os.path.join(d, "potts.pkl"), "wb") as fh:
        pickle.dump({"J_ij": pr["J"], "h_i": pr["h"], "index_list": index_list, "reg_coef": 1.0}, fh)
    for k, net in enumerate(pr["cnn"]):
        sd = {"encoder.weight": torch.from_numpy(net["W0"]), "encoder.bias": torch.from_numpy(net["b0"]),
              "embedding.0.weight": torch.from_numpy(net["W1"]), "embedding.0.bias": torch.from_numpy(net["b1"]),
              "decoder.weight": torch.from_numpy(net["d"])[None], "decoder.bias": torch.from_numpy(net["c"])}
        torch.save({"model": sd}, os.path.join(d, f"onehot_cnn_seed={k}.pt"))
    return d


def build(pr, n_chains, lamda, pas, nmut, paper, workdir=None):
    """-> (energy, sampler, initial_population, min_pos, max_pos) : the reference's own objects on device 'cpu'."""
    PoE, PPDE_PAS = _import_reference()
    root = workdir or tempfile.mkdtemp(prefix="ppde_ref_")
    write_weights_dir(root, pr)
    args = argparse.Namespace(energy_lamda=lamda, unsupervised_expert="potts", protein_weights=root, protein=PROTEIN,
                              n_chains=n_chains, device="cpu", ppde_pas_length=pas, nmut_threshold=nmut, paper_results=paper)
    with contextlib.redirect_stdout(io.StringIO()):
        energy = PoE(args)
    sampler = PPDE_PAS(args)
    pop = energy.wt_onehot.repeat(n_chains, 1, 1)
    lo = int(pr["win_lo"])
    return energy, sampler, pop, lo, lo + int(pr["J"].shape[0]) - 1


def time_steps(pr, n_chains, lamda, pas, nmut, paper, steps, warmup):
    """Steady-state seconds per MCMC iteration of the reference's own loop: `run` is called once for warmup + steps + 1
    iterations with log_every = 1, and the `oracle` callable the loop invokes at every log (ppde.py:156) is used as the
    clock - the t = 0 block, the first `warmup` iterations and the final gathering are outside the interval.
    -> (seconds for `steps` iterations, threads)"""
    energy, sampler, pop, lo, hi = build(pr, n_chains, lamda, pas, nmut, paper)
    stamps = []

    def oracle(x):
        stamps.append(time.perf_counter())
        return torch.zeros(x.shape[0])

    T = warmup + steps + 1
    with contextlib.redirect_stdout(io.StringIO()):
        sampler.run(pop, T, energy, lo, hi, oracle, log_every=1)
    # stamps[0]: t = 0 block; stamps[i] for i >= 1: log of iteration i (after its accept / reset)
    assert len(stamps) == T, (len(stamps), T)
    w = max(warmup, 1)
    return stamps[w + steps] - stamps[w], torch.get_num_threads()
