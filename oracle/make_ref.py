#!/usr/bin/env python
"""ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.  Recipe for `oracle/_ref/`.

The reference's hot path is ONE pure-Python file, `/root/reference/mpvae.py` (no build system, nothing to
compile).  This script places an UNMODIFIED copy of it under `oracle/_ref/` -- git-ignored, so it never enters the
history, but shipped to the GPU box with the snapshot like any other built artefact -- so that

  * `bench.py --impl reference` and the `cpu_baseline` leg can time the reference's own `compute_loss`
    (mpvae.py:145-210) on the GPU box's host cores (`cpu_baseline.kind = "reference"`), and
  * `bench.py`'s `torch_cuda_baseline` can run the same unmodified function with CUDA tensors, i.e. what a user of
    the reference's train.py:20 (`cuda:0`) sees on the same B200.

Nothing in `mpvae-1_b200/` imports it.  `__graft_entry__.build()` runs this when `/root/reference` exists; on the GPU
box the prebuilt copy is used.  The SHA-256 of the source is recorded next to the copy.

    python oracle/make_ref.py [--reference /root/reference]
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ("mpvae.py",)          # the hot path (VAE.forward + compute_loss); evals.py etc. are not needed to time it


def make(reference: str = "/root/reference", quiet: bool = False) -> bool:
    if not os.path.isdir(reference):
        if not quiet:
            print(f"[make_ref] {reference} not present; keeping whatever is in {OUT}")
        return os.path.exists(os.path.join(OUT, FILES[0]))
    os.makedirs(OUT, exist_ok=True)
    lines = []
    for name in FILES:
        src, dst = os.path.join(reference, name), os.path.join(OUT, name)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            lines.append(f"{hashlib.sha256(f.read()).hexdigest()}  {name}  (verbatim copy of {src})")
    with open(os.path.join(OUT, "SOURCE.txt"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if not quiet:
        print("[make_ref] " + "; ".join(lines))
    return True


def load():
    """The unmodified reference module (`oracle/_ref/mpvae.py`), or None when the recipe has not run."""
    path = os.path.join(OUT, "mpvae.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("mpvae_reference_unmodified", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    sys.exit(0 if make(ap.parse_args().reference) else 1)
