#!/usr/bin/env python
"""Error statistics of from_root_positions on the GPU against the float64 reference algorithm (oracle), per skeleton:
quantiles of |rotation - reference|, where the outliers sit (joints with several children?) and the error of the pose
rebuilt from the rotations.  One JSON line per skeleton.

    python tests/dev/frp_error_stats.py [frames_22 frames_52 frames_65]
"""
import json
import os
import sys
import warnings

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
warnings.simplefilter("ignore")

from oracle import pymotion_oracle as orc  # noqa: E402
from pymotion_b200.ops import skeleton as sk  # noqa: E402
from pymotion_b200.topologies import parents_of, synth_numpy  # noqa: E402

sizes = [int(x) for x in sys.argv[1:]] or [20000, 6000, 4000]
names = ("body22", "smplh52", "deep65") * (len(sizes) // 3)
for name, n in zip(names, sizes):
    par = parents_of(name)
    rot, gp, off = synth_numpy(n, par, seed=n)
    pos, _ = orc.fk(rot, gp, off, par)
    centred = (pos - pos[:, 0:1]).astype(np.float32)
    want = orc.from_root_positions(centred.astype(np.float64), par, off.astype(np.float64))
    got = np.asarray(sk.from_root_positions(centred, par, off), dtype=np.float64)
    d = np.abs(got - want).max(axis=-1)  # per (frame, joint)
    n_children = np.bincount(par[1:], minlength=len(par))
    multi = n_children > 1
    zero = np.zeros((1, 3))
    p_got, _ = orc.fk(got, zero, off.astype(np.float64), par)
    p_want, _ = orc.fk(want, zero, off.astype(np.float64), par)
    q = lambda x, p: float(np.quantile(x, p))
    print(json.dumps({
        "skeleton": name, "frames": n,
        "median": q(d, 0.5), "p99": q(d, 0.99), "p999": q(d, 0.999), "p9999": q(d, 0.9999), "max": float(d.max()),
        "frac_above_1e-5": float((d > 1e-5).mean()), "frac_above_1e-4": float((d > 1e-4).mean()),
        "max_single_child_joints": float(d[:, ~multi].max()), "max_multi_child_joints": float(d[:, multi].max()) if multi.any() else 0.0,
        "p999_single_child": q(d[:, ~multi], 0.999),
        "pose_rebuild_max": float(np.abs(p_got - p_want).max()), "pose_vs_input_max": float(np.abs(p_got - centred).max()),
    }), flush=True)
