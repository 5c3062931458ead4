"""The drop-in driver (scripts/directed_evolution.py) end to end on the reference's on-disk formats: flags, files read,
files written (reference scripts/directed_evolution.py:91-101), and the 6-tuple contract (ppde.py:191-192)."""
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_driver_writes_reference_result_files(tmp_path):
    sys.path.insert(0, os.path.join(REPO, "tools")); sys.path.insert(0, os.path.join(REPO, "scripts"))
    import make_weights_dir
    import directed_evolution as drv
    make_weights_dir.write("PABP", str(tmp_path / "weights"), window=(3, 90), offset=115)
    args = drv.build_parser().parse_args([
        "--protein_weights", str(tmp_path / "weights"), "--results_path", str(tmp_path / "results"),
        "--protein", "PABP_YEAST_Fields2013", "--n_chains", "64", "--n_iters", "12", "--log_every", "5",
        "--energy_lamda", "5", "--nmut_threshold", "10", "--seed", "3", "--disable_MSA_transformer_scoring"])
    out = drv.main(args)
    n, T, L = 64, 12, 96
    shapes = {"population.npy": (n, L, 20), "pred_fitness_scores.npy": (n,), "oracle_fitness_scores.npy": (n,),
              "potts_scores.npy": (n,), "energy_scores.npy": (n,), "energy_history.npy": (T + 1, n),
              "fitness_history.npy": (T + 1, n)}
    for f, shp in shapes.items():
        a = np.load(out / f)
        assert a.shape == shp, (f, a.shape)
        assert np.isfinite(a).all()
    pop = np.load(out / "population.npy")
    assert np.array_equal(pop.sum(-1), np.ones((n, L), dtype=pop.dtype))          # one-hot
    cfg = json.load(open(out / "config.txt"))
    assert cfg["nmut_threshold"] == 10 and cfg["n_chains"] == 64
    # best energy = first maximum of the history (ppde.py:173); hard threshold respected by the best samples
    eh = np.load(out / "energy_history.npy")
    assert np.allclose(np.load(out / "energy_scores.npy"), eh.max(0))
    assert (eh[0] == eh[0, 0]).all()                                               # all chains start at the wild type
    from ppde_b200 import weights as W
    wt = W.seq_to_aa(str(np.load(os.path.join(REPO, "tests", "golden", "weights_PABP.npz"))["wt_seq"]))
    assert ((pop.argmax(-1) != wt[None]).sum(1) <= 10 + 3).all()                   # recorded pre-reset: <= thr + path length
