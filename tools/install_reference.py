#!/usr/bin/env python
"""Place an UNMODIFIED copy of the reference under baseline/_ref/ (git-ignored, travels to the GPU box with gpurun).

    python tools/install_reference.py

The reference is a directory of Python scripts without setup.py / pyproject.toml, so `pip install` cannot be used;
a plain tree copy is the install.  Nothing under baseline/_ref is product source: only tests/ (reference-on-CUDA
parity) and bench.py (`--impl reference`, `reference_gpu` leg) import it, always through tests/ref_loader.py.
__graft_entry__.build() calls this when /root/reference exists (the build container); on the GPU box the copy that
travelled with the snapshot is used as is.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("FALN_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def install(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print(f"{SRC} not present: keeping whatever is in {DST}")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    if verbose:
        n = sum(len(f) for _, _, f in os.walk(DST))
        print(f"reference copied to {DST} ({n} files, unmodified)")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
