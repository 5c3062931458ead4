"""The C-ABI library loads and exports every symbol include/ppde_b200.h declares; ctypes mirrors
match the C struct layouts (checked with gcc). No compute calls: runs without a GPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "ppde_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(REPO, "ppde_b200", "libppde_b200.so")):
        g.build()
    from ppde_b200 import _lib
    return _lib.load()


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppde_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from ppde_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ppde_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in ppde_b200/_lib.py"
    assert lib.ppde_version().decode().startswith("ppde_b200")


def test_struct_layouts_match_c(tmp_path):
    from ppde_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "ppde_b200.h"\n'
        'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(ppde_potts_t), sizeof(ppde_cnn_net_t),'
        ' sizeof(ppde_cnn_t), sizeof(ppde_chains_t), sizeof(ppde_pas_params_t), offsetof(ppde_chains_t, traj_chain),'
        ' offsetof(ppde_pas_params_t, t_dev)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), "-o", str(exe), str(prog)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    got = [int(x) for x in out]
    want = [ctypes.sizeof(_lib.PottsT), ctypes.sizeof(_lib.CnnNetT), ctypes.sizeof(_lib.CnnT),
            ctypes.sizeof(_lib.ChainsT), ctypes.sizeof(_lib.PasParamsT), _lib.ChainsT.traj_chain.offset,
            _lib.PasParamsT.t_dev.offset]
    assert got == want


def test_product_fails_loudly_without_cuda():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from ppde_b200.engine import PoEModel
    from ppde_b200.synthetic import synthetic_problem
    pr = synthetic_problem(12)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PoEModel(pr["wt"], pr["J"], pr["h"], pr["win_lo"], pr["cnn"], 1.0)


def test_product_never_imports_the_oracle():
    """The product path must not import, include, load or execute anything under oracle/ (the CPU checker).
    ("oracle" is also the reference's name for its fitness model, ppde/nets.py:315 - that word is fine.)"""
    import re
    pkg = os.path.join(REPO, "ppde_b200")
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|importlib[^\n]*['\"]oracle|#include\s*[\"<][^\n]*oracle/|['\"][^'\"\n]*oracle/[^'\"\n]*['\"]",
                     re.MULTILINE)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                m = bad.search(text)
                assert m is None, f"{f} reaches into the oracle package: {m.group(0)!r}"
