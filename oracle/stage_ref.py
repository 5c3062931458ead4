"""TEST INFRASTRUCTURE — stages the UNMODIFIED reference package for the reference arm of bench.py.

    python -m oracle.stage_ref            (run by __graft_entry__.build() where /root/reference exists)

The reference (pemami4911/ppde) is a pure-Python package whose `pyproject.toml` needs `poetry-core` to build; that backend is not
in this image's wheelhouse, so `pip install --target` fails (recorded in DESIGN.md).  This script does what installing the
wheel would do: it copies the package's `*.py` files, byte for byte, from /root/reference/ppde into the git-ignored directory
oracle/_ref/ppde (it travels to the GPU box like a built .so; it is never committed) and writes a manifest with their SHA-256
digests.  Nothing is patched.  `oracle/ref_arm.py` imports it from there.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC_ROOT = os.environ.get("PPDE_REFERENCE_ROOT", "/root/reference")


def stage(verbose=True):
    src = os.path.join(SRC_ROOT, "ppde")
    if not os.path.isdir(src):
        if verbose:
            print(f"stage_ref: {src} not present; keeping whatever is staged under {DST}")
        return os.path.isdir(os.path.join(DST, "ppde"))
    dst = os.path.join(DST, "ppde")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    manifest = {}
    for root, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if d != "__pycache__"]
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(root, f), src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), out)
            with open(out, "rb") as fh:
                manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
        json.dump({"source": src, "files": manifest}, fh, indent=1, sort_keys=True)
    if verbose:
        print(f"stage_ref: {len(manifest)} files -> {dst}")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
