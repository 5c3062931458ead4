"""CPU-only checks of the host side of the boundary: the host one-hot <-> residue codec of the C-ABI library (no CUDA call), the
staged reference arm (oracle/_ref) and bench.py's reference leg."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ppde_port as port

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n,L,threads", [(1, 7, 1), (33, 96, 3), (5000, 238, 8)])
def test_host_codec_round_trip(n, L, threads):
    from ppde_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n + L)
    aa = rng.integers(0, 20, size=(n, L)).astype(np.uint8)
    x = port.aa_to_onehot(aa).contiguous()
    stride = (L + 15) // 16 * 16
    out = np.full((n, stride), 255, dtype=np.uint8)
    assert lib.ppde_host_onehot_to_aa(x.data_ptr(), n, L, out.ctypes.data, stride, threads) == 0
    assert np.array_equal(out[:, :L], aa) and (out[:, L:] == 0).all()
    x2 = torch.full((n, L, 20), 7.0)
    assert lib.ppde_host_aa_to_onehot(out.ctypes.data, stride, n, L, x2.data_ptr(), threads) == 0
    assert torch.equal(x, x2)
    # first maximum on ties, like torch.argmax (data_utils.onehot2seq)
    t = torch.zeros(1, L, 20); t[0, :, 3] = 1.0; t[0, :, 11] = 1.0
    assert lib.ppde_host_onehot_to_aa(t.data_ptr(), 1, L, out.ctypes.data, stride, 1) == 0
    assert (out[0, :L] == 3).all()
    assert lib.ppde_host_onehot_to_aa(None, 1, L, out.ctypes.data, stride, 1) != 0          # NULL -> cudaErrorInvalidValue


def test_reference_arm_runs_the_staged_reference():
    """oracle/_ref holds byte-identical copies of the reference's package files (manifest of SHA-256 digests) and the arm
    drives its own PPDE_PAS.run / ProteinProductOfExperts on weights written in the reference's on-disk formats."""
    from oracle import ref_arm, stage_ref
    if not ref_arm.available():
        if not stage_ref.stage(verbose=False):
            pytest.skip("reference not staged and /root/reference absent")
    import hashlib
    man = json.load(open(os.path.join(ref_arm.REF_DIR, "MANIFEST.json")))
    for rel, digest in man["files"].items():
        with open(os.path.join(ref_arm.REF_DIR, "ppde", rel), "rb") as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == digest, rel
        src = os.path.join("/root/reference/ppde", rel)
        if os.path.exists(src):
            with open(src, "rb") as fh:
                assert hashlib.sha256(fh.read()).hexdigest() == digest, f"{rel} differs from the reference checkout"
    from ppde_b200.synthetic import synthetic_problem
    pr = synthetic_problem(40, seed=3, window=(2, 36))
    energy, sampler, pop, lo, hi = ref_arm.build(pr, 6, 2.0, 2, 3, False)
    assert (lo, hi) == (2, 36) and tuple(pop.shape) == (6, 40, 20)
    # the reference's energy on these files == the port's on the same arrays (the port is what the CUDA path is tested against)
    w = port.Weights(wt=pr["wt"], J=pr["J"], h=pr["h"], win_lo=pr["win_lo"], cnn=pr["cnn"], lamda=2.0)
    x = port.aa_to_onehot(np.random.default_rng(0).integers(0, 20, size=(5, 40)))
    e_ref, f_ref, g_ref = energy.get_energy_and_grads(x.clone().requires_grad_())
    e_p, f_p, g_p = port.PortEnergy(w).get_energy_and_grads(x)
    assert torch.allclose(e_ref.detach(), e_p, rtol=1e-6, atol=1e-6) and torch.allclose(g_ref, g_p, rtol=1e-6, atol=1e-6)
    dt, threads = ref_arm.time_steps(pr, 6, 2.0, 2, 3, False, steps=2, warmup=1)
    assert dt > 0 and threads >= 1


def test_bench_reference_leg_prints_the_contract_line():
    env = dict(os.environ, PYTHONPATH=REPO)
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--workload", "pabp_readme_128",
                        "--cpu-chains", "8", "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "chain-steps/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
