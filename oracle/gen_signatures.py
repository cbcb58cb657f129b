#!/usr/bin/env python
"""Record the public function signatures of the reference modules this repository mirrors
(run in the build container: needs /root/reference).  Output: tests/golden/signatures.json, compared with the
drop-in package by tests/test_signatures.py.  TEST INFRASTRUCTURE ONLY."""
import importlib
import inspect
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")

MODULES = ["ops.skeleton", "ops.center_of_mass", "ops.time", "ops.vector", "rotations.quat", "rotations.dual_quat",
           "rotations.ortho6d"]


def describe(fn):
    out = []
    for name, p in inspect.signature(fn).parameters.items():
        default = None if p.default is inspect.Parameter.empty else repr(p.default)
        out.append([name, default])
    return out


def main():
    table = {}
    for mod in MODULES:
        for twin in ("", "_torch"):
            m = importlib.import_module(f"pymotion.{mod}{twin}")
            funcs = {n: describe(f) for n, f in vars(m).items()
                     if inspect.isfunction(f) and f.__module__ == m.__name__ and not n.startswith("_")}
            table[f"{mod}{twin}"] = funcs
    path = os.path.join(REPO, "tests", "golden", "signatures.json")
    with open(path, "w") as fh:
        json.dump(table, fh, indent=1, sort_keys=True)
    print(path, {k: len(v) for k, v in table.items()})


if __name__ == "__main__":
    main()
