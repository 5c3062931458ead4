"""Readers for the reference's on-disk formats (no biopython needed).

  wt.fasta                       ppde/third_party/hsu/io_utils.py:178-188; id `NAME/start-end` gives
                                 the Potts index offset (ppde/nets.py:257-261)
  potts.pkl                      dict J_ij [Lp,Lp,20,20], h_i [Lp,20], index_list [Lp], reg_coef
                                 (ppde/nets.py:247-251)
  onehot_cnn_seed={0,1,2}.pt     torch checkpoint, state dict under ['model'] (ppde/nets.py:417-424)
  results-...-seed={0..19}-linear.pkl   ridge heads of the oracle model (ppde/nets.py:323-329)
"""
import os
import pickle

import numpy as np

ALPHABET = "ACDEFGHIKLMNPQRSTVWY"          # data_utils.py:48-72: A=0 ... Y=19
_AA_TO_INT = {c: i for i, c in enumerate(ALPHABET)}


def read_fasta(path):
    """-> (list of sequences, list of ids); same record semantics as Bio.SeqIO.parse(..., 'fasta')."""
    seqs, ids, cur = [], [], None
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                ids.append(line[1:].split()[0] if len(line) > 1 else "")
                seqs.append([])
                cur = seqs[-1]
            elif cur is not None:
                cur.append(line)
    return ["".join(s) for s in seqs], ids


def seq_to_aa(seq):
    try:
        return np.array([_AA_TO_INT[c] for c in seq], dtype=np.uint8)
    except KeyError as e:
        raise ValueError(f"residue {e} is not one of {ALPHABET}") from None


def aa_to_seq(aa):
    return "".join(ALPHABET[int(a)] for a in aa)


def fasta_offset(fasta_id):
    return int(fasta_id.split("/")[-1].split("-")[0]) if "/" in fasta_id else 1


def load_potts(dataset_dir, fasta_id):
    """-> dict(J, h, win_lo, win_hi, reg_coef). index_list must be contiguous (the reference slices
    x[:, index_list[0]:index_list[-1]+1], ppde/nets.py:280)."""
    with open(os.path.join(dataset_dir, "potts.pkl"), "rb") as fh:
        p = pickle.load(fh)
    idx = np.asarray(p["index_list"]).astype(np.int64) - fasta_offset(fasta_id)
    J = np.asarray(p["J_ij"], dtype=np.float32)
    h = np.asarray(p["h_i"], dtype=np.float32)
    if idx[-1] - idx[0] + 1 != J.shape[0]:
        raise ValueError("potts.pkl index_list is not contiguous; the reference's window slice would not match J")
    return {"J": J, "h": h, "win_lo": int(idx[0]), "win_hi": int(idx[-1]), "reg_coef": float(p["reg_coef"])}


def cnn_from_state_dict(sd):
    """OnehotCNN state dict (ppde/nets.py:350-361) -> plain arrays."""
    g = lambda k: np.asarray(sd[k].detach().cpu().numpy() if hasattr(sd[k], "detach") else sd[k], dtype=np.float32)
    return {"W0": g("encoder.weight"), "b0": g("encoder.bias"), "W1": g("embedding.0.weight"),
            "b1": g("embedding.0.bias"), "d": g("decoder.weight").reshape(-1), "c": g("decoder.bias").reshape(-1)}


def load_cnn_ensemble(dataset_dir, n=3):
    import torch
    nets = []
    for k in range(n):
        ck = torch.load(os.path.join(dataset_dir, f"onehot_cnn_seed={k}.pt"), map_location="cpu")
        nets.append(cnn_from_state_dict(ck["model"]))
    return nets


def load_ridge_heads(dataset_dir, n=20):
    """-> (coef f32 [n, 1+20L], intercept f32 [n], reg f64 [n]) cast as the reference does (nets.py:327-329)."""
    coefs, icpt, regs = [], [], []
    for seed in range(n):
        with open(os.path.join(dataset_dir, f"results-predictor=ev+onehot-train=-1-seed={seed}-linear.pkl"), "rb") as fh:
            r = pickle.load(fh)
        coefs.append(np.asarray(r["coef_"]).astype(np.float32))
        icpt.append(np.float32(r["intercept_"]))
        regs.append(float(r["reg_coef"]))
    return np.stack(coefs), np.asarray(icpt, dtype=np.float32), np.asarray(regs, dtype=np.float64)
