"""TEST INFRASTRUCTURE — drives the UNMODIFIED reference (pemami4911/ppde) in this container.

Only usable where /root/reference exists (the build container); nothing under
tests/ -m gpu, smoke() or bench.py imports this module.  It is the tool that
generates the committed golden vectors (oracle/make_golden.py) and pins the
restatement in oracle/ppde_port.py against the reference's own code.

What it provides:
  * sys.modules shims for the two un-installed imports the reference needs
    (`Bio.SeqIO.parse`, used at ppde/third_party/hsu/io_utils.py:5,178-188, and
    `esm_one_hot.pretrained`, imported at ppde/nets.py:11 but unused on our path);
  * a synthetic `potts.pkl` writer (the real ones are missing from the checkout,
    .MISSING_LARGE_BLOBS:3-5) with the key layout ppde/nets.py:247-262 reads;
  * a scratch weights directory that symlinks the shipped `wt.fasta`, CNN
    checkpoints and ridge pickles next to the synthetic Potts file;
  * `SharedStreams`: patches `torch.randint`, `torch.multinomial`,
    `torch.rand_like` so that the reference consumes the indexed Philox streams
    of ppde_b200/philox.py instead of its global generators.
"""
import argparse
import os
import pickle
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("PPDE_REFERENCE_ROOT", "/root/reference")

PROTEINS = {
    "PABP": "PABP_YEAST_Fields2013",
    "UBE4B": "UBE4B_MOUSE_Klevit2013-nscor_log2_ratio",
    "GFP": "GFP_AEQVI_Sarkisyan2016",
}


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ppde"))


def install_shims():
    """Make `import ppde.energy` work without biopython / esm_one_hot."""
    if "Bio" not in sys.modules:
        bio = types.ModuleType("Bio")
        seqio = types.ModuleType("Bio.SeqIO")

        class _Record:
            def __init__(self, rid, seq):
                self.id, self.seq = rid, seq

        def parse(filename, fmt):
            assert fmt == "fasta"
            rid, chunks = None, []
            with open(filename) as fh:
                for line in fh:
                    line = line.strip()
                    if not line:
                        continue
                    if line.startswith(">"):
                        if rid is not None:
                            yield _Record(rid, "".join(chunks))
                        rid, chunks = line[1:].split()[0], []
                    else:
                        chunks.append(line)
            if rid is not None:
                yield _Record(rid, "".join(chunks))

        seqio.parse = parse
        bio.SeqIO = seqio
        sys.modules["Bio"] = bio
        sys.modules["Bio.SeqIO"] = seqio
    if "esm_one_hot" not in sys.modules:
        esm = types.ModuleType("esm_one_hot")
        pre = types.ModuleType("esm_one_hot.pretrained")
        esm.pretrained = pre
        sys.modules["esm_one_hot"] = esm
        sys.modules["esm_one_hot.pretrained"] = pre
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def fasta_offset(fasta_path):
    """Offset rule of ppde/nets.py:257-261 (id `NAME/start-end` -> start, else 1)."""
    with open(fasta_path) as fh:
        rid = fh.readline().strip()[1:].split()[0]
    return int(rid.split("/")[-1].split("-")[0]) if "/" in rid else 1


from oracle.ppde_port import synthetic_potts  # noqa: E402  (one definition, shared with the port)


def make_weights_dir(dst_root, protein_key, window=None, seed=0, symmetric=True,
                     zero_diag=True, sigma_j=0.05, sigma_h=0.5):
    """Scratch copy of weights/<protein>/ with a synthetic potts.pkl. Returns (root, protein)."""
    protein = PROTEINS[protein_key]
    src = os.path.join(REFERENCE_ROOT, "weights", protein)
    dst = os.path.join(dst_root, protein)
    os.makedirs(dst, exist_ok=True)
    for name in os.listdir(src):
        link = os.path.join(dst, name)
        if not os.path.lexists(link):
            os.symlink(os.path.join(src, name), link)
    fasta = os.path.join(src, "wt.fasta")
    with open(fasta) as fh:
        L = len("".join(l.strip() for l in fh.readlines()[1:]))
    lo, hi = window if window is not None else (0, L - 1)
    Lp = hi - lo + 1
    J, h = synthetic_potts(Lp, seed, sigma_j, sigma_h, symmetric, zero_diag)
    index_list = np.arange(lo, hi + 1, dtype=np.int64) + fasta_offset(fasta)
    with open(os.path.join(dst, "potts.pkl"), "wb") as fh:
        pickle.dump({"J_ij": J, "h_i": h, "index_list": index_list, "reg_coef": 1.0}, fh)
    return dst_root, protein


def make_args(weights_root, protein, n_chains, lamda, pas=2, nmut=0, paper=False):
    """Namespace with the attributes the reference classes read
    (ppde/energy.py:72-95, ppde/protein_samplers/ppde.py:9-17)."""
    return argparse.Namespace(
        energy_lamda=lamda, unsupervised_expert="potts", protein_weights=weights_root,
        protein=protein, n_chains=n_chains, device="cpu", ppde_pas_length=pas,
        nmut_threshold=nmut, paper_results=paper)


class SharedStreams:
    """Context manager: the reference draws from the indexed Philox streams.

    Call order inside `PPDE_PAS.run` (ppde.py:65-139) per iteration t:
    one `torch.randint` (path lengths), `max_u` x `torch.multinomial`
    (sub-steps s = 0..max_u-1), one `torch.rand_like` (accept).
    `torch.multinomial(p, 1, True)` is `argmax(p / Exp(1))`;
    the exponentials are `-log(u)` of the shared uniforms.
    """

    def __init__(self, seed, n_chains, chain_offset=0, record=None):
        from ppde_b200 import philox
        self.philox = philox
        self.seed = seed
        self.chains = np.arange(n_chains, dtype=np.uint32) + np.uint32(chain_offset)
        self.t = -1
        self.s = 0
        self.record = record if record is not None else {}
        for k in ("U", "idx", "u_acc", "p_at_idx"):
            self.record.setdefault(k, [])

    def __enter__(self):
        self._orig = (torch.randint, torch.multinomial, torch.rand_like)
        me = self

        def randint(low, high, size=None, **kw):
            me.t += 1
            me.s = 0
            pas2 = high  # high = 2*pas (exclusive)
            U = me.philox.path_lengths(me.seed, me.t, me.chains, pas2 // 2)
            assert low == 1 and U.max() < high
            me.record["U"].append(U.copy())
            me.record["idx"].append([])
            me.record["p_at_idx"].append([])
            return torch.from_numpy(U.astype(np.int64)).reshape(size)

        def multinomial(p, num_samples, replacement=False, **kw):
            assert num_samples == 1
            u = me.philox.proposal_uniforms(me.seed, me.t, me.s, me.chains, p.shape[-1])
            e = -torch.log(torch.from_numpy(u))
            idx = torch.argmax(p / e, dim=-1, keepdim=True)
            me.record["idx"][-1].append(idx[:, 0].numpy().copy())
            me.record["p_at_idx"][-1].append(p.gather(-1, idx)[:, 0].numpy().copy())
            me.s += 1
            return idx

        def rand_like(x, **kw):
            u = me.philox.accept_uniforms(me.seed, me.t, me.chains)
            me.record["u_acc"].append(u.copy())
            return torch.from_numpy(u).reshape(x.shape).to(x.dtype)

        torch.randint, torch.multinomial, torch.rand_like = randint, multinomial, rand_like
        return self

    def __exit__(self, *exc):
        torch.randint, torch.multinomial, torch.rand_like = self._orig
        return False


class EnergySpy:
    """Wraps a reference energy object; records every get_energy_and_grads call
    (inputs as residue indices, outputs) without touching reference code."""

    def __init__(self, energy, keep_grads=False):
        self._e = energy
        self.calls = []
        self.keep_grads = keep_grads
        self.wt_onehot = energy.wt_onehot
        self.lamda = getattr(energy, "lamda", None)

    def get_energy(self, x):
        return self._e.get_energy(x)

    def get_energy_and_grads(self, x):
        e, f, g = self._e.get_energy_and_grads(x)
        rec = {"aa": x.detach().argmax(-1).numpy().astype(np.uint8),
               "e": e.detach().numpy().copy(), "fit": f.detach().numpy().copy()}
        if self.keep_grads:
            rec["grad"] = g.detach().numpy().copy()
        self.calls.append(rec)
        return e, f, g


def load_reference(weights_root, protein, n_chains, lamda, pas=2, nmut=0, paper=False):
    """Build the reference's energy, sampler and oracle objects (stdout silenced)."""
    install_shims()
    import contextlib
    import io
    from ppde.energy import ProteinProductOfExperts
    from ppde.nets import AugmentedLinearRegression
    from ppde.protein_samplers.ppde import PPDE_PAS
    args = make_args(weights_root, protein, n_chains, lamda, pas, nmut, paper)
    with contextlib.redirect_stdout(io.StringIO()):
        energy = ProteinProductOfExperts(args)
        oracle = AugmentedLinearRegression(os.path.join(weights_root, protein))
    sampler = PPDE_PAS(args)
    return args, energy, sampler, oracle
