#!/usr/bin/env python
"""How often does fk_quat's sign differ from quat.from_matrix(fk(...)[1]) -- the convention it claims (quat.py:85-156)?
Per skeleton: flip rate against (a) this package's own from_matrix of its own fk matrices (fp32, same branches as the
reference) and (b) the float64 oracle on a sample, and how close to a branch tie of from_matrix the flipped entries sit."""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

from oracle import pymotion_oracle as orc  # noqa: E402
from pymotion_b200.ops import skeleton as sk  # noqa: E402
from pymotion_b200.rotations import quat  # noqa: E402
from pymotion_b200.topologies import parents_of, synth_torch  # noqa: E402

dev = torch.device("cuda", 0)
for name, frames in (("body22", 1_000_000), ("smplh52", 400_000), ("deep65", 300_000)):
    par = parents_of(name)
    rot, gp, off = synth_torch(frames, par, dev, seed=7)
    pos, grot = sk.fk_quat(rot, gp, off, par)
    _, rotm = sk.fk(rot, gp, off, par)
    ref32 = quat.from_matrix(rotm)
    dots = (grot * ref32).sum(-1)
    flipped = dots < 0
    m = rotm
    # margins of the three branch tests of from_matrix (quat.py:111-155): m22 < 0 ; m00 > m11 ; m00 < -m11
    t1, t2, t3 = m[..., 2, 2].abs(), (m[..., 0, 0] - m[..., 1, 1]).abs(), (m[..., 0, 0] + m[..., 1, 1]).abs()
    margin = torch.minimum(t1, torch.where(m[..., 2, 2] < 0, t2, t3))
    n_s = 20_000
    r64, g64, o64 = rot[:n_s].cpu().numpy().astype(np.float64), gp[:n_s].cpu().numpy().astype(np.float64), off.cpu().numpy().astype(np.float64)
    want = orc.quat_from_matrix(orc.fk(r64, g64, o64, par)[1])
    d64 = (grot[:n_s].cpu().numpy().astype(np.float64) * want).sum(-1)
    out = {"skeleton": name, "frames": frames, "entries": int(dots.numel()),
           "flips_vs_own_from_matrix": int(flipped.sum()), "flip_rate_vs_own_from_matrix": float(flipped.float().mean()),
           "max_tie_margin_of_flipped": float(margin[flipped].max()) if flipped.any() else 0.0,
           "median_tie_margin_of_flipped": float(margin[flipped].median()) if flipped.any() else 0.0,
           "flip_rate_vs_float64_oracle_20k": float((d64 < 0).mean()), "min_abs_dot": float(dots.abs().min())}
    print(json.dumps(out), flush=True)
    del rot, gp, pos, grot, rotm, ref32, dots, m
    torch.cuda.empty_cache()
