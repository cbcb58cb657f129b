"""Builds pymotion_b200/libpymotion_b200.so in-tree with nvcc for sm_100a.

    python -m pymotion_b200._build [--force] [--verbose]

One object per translation unit (csrc/api_*.cu), compiled in parallel, linked into one shared library.
The .so is git-ignored but travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")
LIB = os.path.join(PKG, "libpymotion_b200.so")
OBJ_DIR = os.path.join(PKG, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC,-O2,-Wall"]


def units() -> list[str]:
    return sorted(os.path.join(CSRC, n) for n in os.listdir(CSRC) if n.startswith("api_") and n.endswith(".cu"))


def headers() -> list[str]:
    out = [os.path.join(INCLUDE, "pymotion_b200.h")]
    out += sorted(os.path.join(CSRC, n) for n in os.listdir(CSRC) if n.endswith((".cuh", ".h")))
    return out


def _obj(unit: str) -> str:
    return os.path.join(OBJ_DIR, os.path.basename(unit)[:-3] + ".o")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > built for s in units() + headers())


def _compile(unit: str, force: bool, verbose: bool, newest_header: float) -> str:
    obj = _obj(unit)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(unit), newest_header):
        return obj
    cmd = [NVCC, *ARCH, *FLAGS, "-c", "-o", obj, unit]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed on {os.path.basename(unit)}")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in headers())
    with ThreadPoolExecutor(max_workers=max(1, min(len(units()), os.cpu_count() or 1))) as pool:
        objs = list(pool.map(lambda u: _compile(u, force, verbose, newest_header), units()))
    cmd = [NVCC, *ARCH, "-shared", "--cudart", "static", "-Xcompiler", "-fPIC", "-o", LIB, *objs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libpymotion_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
