#!/usr/bin/env python
"""The reference's OWN float32 envelope for from_root_positions: how far its torch twin (float32,
pymotion/ops/skeleton_torch.py) lands from its NumPy path (float64, ops/skeleton.py:96-170) on the batches the GPU
parity test uses.  Writes tests/golden/ik_twin_envelope.json (run in the build container: needs /root/reference).

The op is ill-conditioned by construction -- np.isclose snaps small alignments to the identity (a frame on the other side
of the threshold differs by up to 2.2e-3 and takes its descendants with it), the roll of a joint with several children
has an arbitrary sign when its axis is perpendicular to the correction -- so "how close is close enough" is answered by
the reference itself: the CUDA kernel is held to a small multiple of these numbers (tests/test_gpu_ik.py).
TEST INFRASTRUCTURE ONLY."""
import json
import os
import sys
import warnings

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")

import pymotion.ops.skeleton as ref_np  # noqa: E402
import pymotion.ops.skeleton_torch as ref_t  # noqa: E402

from pymotion_b200.topologies import parents_of, synth_numpy  # noqa: E402

BATCHES = (("body22", 3001), ("smplh52", 1000), ("deep65", 517), ("body22", 20000), ("smplh52", 6000), ("deep65", 4000))


def batch(name, n):
    """The inputs of tests/test_gpu_ik.py::test_from_root_positions_vs_oracle (seed = frame count)."""
    par = parents_of(name)
    rot, gp, off = synth_numpy(n, par, seed=n)
    pos, _ = ref_np.fk(rot, gp, off, par)
    return par, off, (pos - pos[:, 0:1]).astype(np.float32)


def stats(got, want, par, off):
    d = np.abs(np.asarray(got, dtype=np.float64) - want).max(axis=-1)
    zero = np.zeros((1, 3))
    p_got, _ = ref_np.fk(np.asarray(got, dtype=np.float64), zero, off.astype(np.float64), par)
    p_want, _ = ref_np.fk(want, zero, off.astype(np.float64), par)
    q = lambda p: float(np.quantile(d, p))  # noqa: E731
    return {"median": q(0.5), "p99": q(0.99), "p999": q(0.999), "max": float(d.max()), "pose_rebuild_max": float(np.abs(p_got - p_want).max())}


if __name__ == "__main__":
    out = {}
    for name, n in BATCHES:
        par, off, centred = batch(name, n)
        want = ref_np.from_root_positions(centred.astype(np.float64), par, off.astype(np.float64))
        twin = ref_t.from_root_positions(torch.from_numpy(centred), torch.from_numpy(par), torch.from_numpy(off)).numpy()
        out[f"{name}/{n}"] = stats(twin, want, par, off)
        print(name, n, out[f"{name}/{n}"], flush=True)
    with open(os.path.join(REPO, "tests", "golden", "ik_twin_envelope.json"), "w") as fh:
        json.dump(out, fh, indent=1)
