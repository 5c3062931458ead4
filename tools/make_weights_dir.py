#!/usr/bin/env python
"""Write a `<protein_weights>/<protein>/` directory in the REFERENCE's on-disk layout (ppde/nets.py:247-262, 323-329,
417-424) from the committed fixtures (tests/golden/weights_<PROT>.npz: shipped WT, CNN checkpoints, ridge heads) plus a
synthetic potts.pkl (the real ones are not in the reference checkout).  Used by the driver test and for demos.

usage: python tools/make_weights_dir.py PABP <out_root> [--window lo hi]"""
import argparse
import os
import pickle
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from ppde_b200.synthetic import synthetic_potts        # noqa: E402

NAMES = {"PABP": "PABP_YEAST_Fields2013", "UBE4B": "UBE4B_MOUSE_Klevit2013-nscor_log2_ratio", "GFP": "GFP_AEQVI_Sarkisyan2016"}


def write(prot, out_root, window=None, offset=1):
    z = np.load(os.path.join(REPO, "tests", "golden", f"weights_{prot}.npz"))
    seq = str(z["wt_seq"])
    L = len(seq)
    lo, hi = window if window else (0, L - 1)
    d = os.path.join(out_root, NAMES[prot])
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "wt.fasta"), "w") as fh:
        fh.write(f">{prot}/{offset}-{offset + L - 1}\n{seq}\n")
    J, h = synthetic_potts(hi - lo + 1, seed=0)
    with open(os.path.join(d, "potts.pkl"), "wb") as fh:
        pickle.dump({"J_ij": J, "h_i": h, "index_list": np.arange(lo, hi + 1) + offset, "reg_coef": 1.0}, fh)
    for k in range(3):
        sd = {"encoder.weight": torch.from_numpy(z[f"cnn{k}_W0"]), "encoder.bias": torch.from_numpy(z[f"cnn{k}_b0"]),
              "embedding.0.weight": torch.from_numpy(z[f"cnn{k}_W1"]), "embedding.0.bias": torch.from_numpy(z[f"cnn{k}_b1"]),
              "decoder.weight": torch.from_numpy(z[f"cnn{k}_d"]).reshape(1, -1), "decoder.bias": torch.from_numpy(z[f"cnn{k}_c"]).reshape(-1)}
        torch.save({"model": sd}, os.path.join(d, f"onehot_cnn_seed={k}.pt"))
    for s in range(20):
        with open(os.path.join(d, f"results-predictor=ev+onehot-train=-1-seed={s}-linear.pkl"), "wb") as fh:
            pickle.dump({"coef_": z["ridge_coef"][s].astype(np.float64), "intercept_": float(z["ridge_intercept"][s]),
                         "reg_coef": float(z["ridge_reg"][s])}, fh)
    return d


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("prot", choices=sorted(NAMES))
    ap.add_argument("out_root")
    ap.add_argument("--window", type=int, nargs=2, default=None)
    a = ap.parse_args()
    print(write(a.prot, a.out_root, tuple(a.window) if a.window else None))
