"""Synthetic skeleton topologies and input generators used by the parity tests
and by ``bench.py``.

pymotion convention: ``parents[0] == 0`` (the root is its own parent,
``/root/reference/pymotion/io/bvh.py:124``) and ``parents[i] < i`` (BVH depth
first order, ``io/bvh.py:77-85``).
"""
from __future__ import annotations

import numpy as np

# 22 joints: Hips -> {left leg x4, right leg x4, spine x3 -> {neck, head}, {left arm x4}, {right arm x4}}
# (topology of the example skeleton whose joint names README.md:49 of the reference lists); depth 7.
BODY22 = [0, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 12, 11, 14, 15, 16, 11, 18, 19, 20]

# SMPL-style 22-joint body used as the trunk of the two larger skeletons.
_SMPL_BODY22 = [0, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19]


def _hand(wrist: int, base: int) -> list[int]:
    """Five 3-joint finger chains hanging off ``wrist``; first finger joint index is ``base``."""
    out: list[int] = []
    for finger in range(5):
        out += [wrist, base + 3 * finger, base + 3 * finger + 1]
    return out


# 52 joints (SMPL-H like): body + 2 x 15 finger joints; depth 10.
SMPLH52 = _SMPL_BODY22 + _hand(20, 22) + _hand(21, 37)

# 65 joints (SMPL-X like, deep): body + jaw/eyes on the head + hands + 10 fingertip leaves; depth 11.
DEEP65 = (
    _SMPL_BODY22
    + [15, 15, 15]
    + _hand(20, 25)
    + _hand(21, 40)
    + [27, 30, 33, 36, 39, 42, 45, 48, 51, 54]
)

# Mid-size skeletons used to calibrate the launch heuristics between the bench sizes: the SMPL body with one
# joint per finger (32 joints) and with three 3-joint fingers per hand (40 joints).
BODY32 = _SMPL_BODY22 + [20] * 5 + [21] * 5
BODY40 = _SMPL_BODY22 + _hand(20, 22)[:9] + _hand(21, 31)[:9]

BODY16 = _SMPL_BODY22[:16]
BODY24 = _SMPL_BODY22 + [20, 21]

TOPOLOGIES = {"body22": BODY22, "smplh52": SMPLH52, "deep65": DEEP65, "chain3": [0, 0, 1], "body32": BODY32,
              "body40": BODY40, "body16": BODY16, "body24": BODY24}


def parents_of(name: str) -> np.ndarray:
    return np.asarray(TOPOLOGIES[name], dtype=np.int64)


def depth_of(parents) -> int:
    depth = [0] * len(parents)
    for i in range(1, len(parents)):
        depth[i] = depth[int(parents[i])] + 1
    return max(depth) if depth else 0


def synth_numpy(n_frames: int, parents, seed: int = 0, dtype=np.float32):
    """CPU synthetic batch: unit quaternions uniform on S^3,
    root positions ~ N(0,1), shared offsets ~ N(0, 0.15^2) with offsets[0] = 0."""
    rng = np.random.default_rng(seed)
    n_joints = len(parents)
    rot = rng.standard_normal((n_frames, n_joints, 4))
    rot /= np.linalg.norm(rot, axis=-1, keepdims=True)
    gpos = rng.standard_normal((n_frames, 3))
    offsets = rng.standard_normal((n_joints, 3)) * 0.15
    offsets[0] = 0.0
    return rot.astype(dtype), gpos.astype(dtype), offsets.astype(dtype)


def synth_torch(n_frames: int, parents, device, seed: int = 1234):
    """Device-side synthetic batch for the large configurations (generated on
    the GPU so a 4M x 65 batch never crosses PCIe)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    n_joints = len(parents)
    rot = torch.randn((n_frames, n_joints, 4), device=device, dtype=torch.float32, generator=gen)
    rot /= rot.norm(dim=-1, keepdim=True)
    gpos = torch.randn((n_frames, 3), device=device, dtype=torch.float32, generator=gen)
    offsets = torch.randn((n_joints, 3), device=device, dtype=torch.float32, generator=gen) * 0.15
    offsets[0] = 0.0
    return rot, gpos, offsets
