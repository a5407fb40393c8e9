#!/usr/bin/env python
"""Stages the UNMODIFIED reference metric for the CPU arm of bench.py -- TEST / MEASUREMENT INFRASTRUCTURE.

The reference's metric path is 24 NumPy-only lines (/root/reference/lib/metric.py, thuml/HashGAN).  /root/reference does not
exist on the GPU box, so this recipe copies that one file (plus the empty lib/__init__.py) byte for byte into oracle/_ref/lib/
-- git-ignored (never part of the repo history) but not gpurun-ignored, so it travels to the box like the built .so files.
`bench.py --impl reference` and bench.py's cpu_baseline leg import MAPs from there (kind: "reference") and fall back to the
restatement oracle/maps_oracle.py (kind: "port") only when the staged copy is missing.  __graft_entry__.build() runs this in
the build container; the product (hashgan_b200/) never touches oracle/_ref.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HASHGAN_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def stage() -> str | None:
    """Copies lib/metric.py from the mounted reference; returns the staged path (None when the reference is not mounted and
    nothing was staged before)."""
    src = os.path.join(REF, "lib", "metric.py")
    dst = os.path.join(DST, "lib", "metric.py")
    if os.path.exists(src):
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        open(os.path.join(DST, "lib", "__init__.py"), "w").close()
        with open(os.path.join(DST, "SOURCE.txt"), "w") as fh:
            fh.write(f"lib/metric.py copied unmodified from {src}\nsha256 {hashlib.sha256(open(src, 'rb').read()).hexdigest()}\n")
    return dst if os.path.exists(dst) else None


def reference_maps_class():
    """The reference's own MAPs class from the staged file (None when it was never staged)."""
    path = os.path.join(DST, "lib", "metric.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("_hashgan_reference_metric", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MAPs


if __name__ == "__main__":
    p = stage()
    print("staged" if p else "reference not mounted; nothing staged", p or "")
    sys.exit(0)
