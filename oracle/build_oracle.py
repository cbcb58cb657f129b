#!/usr/bin/env python
"""Compile oracle/oracle_c.c -> oracle/_build/liboracle_c.so (gcc, OpenMP).
TEST INFRASTRUCTURE ONLY; called by __graft_entry__.build() and lazily by
oracle/oracle_c.py.  The reference itself is pure Python, so there is no
`oracle/_ref` native build: the real reference is exercised in the build
container by oracle/gen_golden.py instead."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle_c.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle_c.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-fno-fast-math",
           "-Wall", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
