#!/usr/bin/env python
"""Stage the UNMODIFIED reference where the GPU box can import it: oracle/_ref/pymotion.

    python oracle/fetch_ref.py            (also called by __graft_entry__.build())

/root/reference exists only in the build container; oracle/_ref/ is git-ignored (no reference source ever
enters the history) but NOT gpurun-ignored, so the copy travels with the snapshot like a built .so.  It is
TEST / BENCH INFRASTRUCTURE ONLY: `bench.py --impl reference` and the `cpu_baseline` leg time it
(`kind: "reference"`), nothing under pymotion_b200/ imports it.  The reference is pure Python + NumPy
(pyproject.toml:23-25), so "building" it is a file copy of the numeric package; its tests, viewer assets and
Blender bridge are left behind.
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/pymotion"
DST_ROOT = os.path.join(HERE, "_ref")
DST = os.path.join(DST_ROOT, "pymotion")
SKIP_DIRS = {"tests", "render", "__pycache__"}


def fetch(force: bool = False) -> str | None:
    """Copy the reference package if the source tree is present; returns the import root or None."""
    if not os.path.isdir(SRC):
        return DST_ROOT if os.path.isdir(DST) else None
    if os.path.isdir(DST) and not force:
        newest_src = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(SRC) for f in fs)
        stamp = os.path.join(DST_ROOT, ".stamp")
        if os.path.exists(stamp) and os.path.getmtime(stamp) >= newest_src:
            return DST_ROOT
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST_ROOT, exist_ok=True)
    shutil.copytree(SRC, DST, ignore=lambda d, names: [n for n in names if n in SKIP_DIRS])
    with open(os.path.join(DST_ROOT, ".stamp"), "w") as fh:
        fh.write("copied from /root/reference/pymotion (unmodified)\n")
    return DST_ROOT


def import_reference():
    """(skeleton, quat, dual_quat) modules of the real reference, or None when oracle/_ref is absent."""
    root = fetch()
    if root is None:
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    import importlib

    return tuple(importlib.import_module(m) for m in
                 ("pymotion.ops.skeleton", "pymotion.rotations.quat", "pymotion.rotations.dual_quat"))


if __name__ == "__main__":
    print(fetch(force="--force" in sys.argv))
