"""The NUMERIC part of ``pymotion.io.bvh.BVH.get_data`` (/root/reference/pymotion/io/bvh.py:332-365) on the GPU.

``BVH.get_data`` turns the Euler angles of a parsed file into the quaternions every op of the fk path consumes::

    rots = quat.normalize(quat.unroll(quat.from_euler(np.radians(rotations), order tiled over frames), axis=0))

Here that chain is ONE fused scan (``pmb_bvh_rotations_to_quat_f32``): the un-unrolled quaternions are never
written, 42 bytes of traffic per entry instead of 110 for the three separate ops.  Parsing the BVH text stays with
the reference's host code (out of scope): pass its ``BVH.data`` dictionary, or the arrays.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _runtime as rt

_AXES = {"x": 0, "y": 1, "z": 2}


def _order_codes_per_joint(rot_order, n_joints: int, device) -> torch.Tensor:
    arr = np.asarray(rot_order)
    if arr.shape != (n_joints, 3):
        raise ValueError(f"rot_order must have shape [{n_joints}, 3], got {arr.shape}")
    idx = np.full(arr.shape, -1, dtype=np.int64)
    for ch, k in _AXES.items():
        idx[arr == ch] = k
    if (idx < 0).any():
        raise KeyError("rot_order entries must be 'x', 'y' or 'z'")  # the reference's dict lookup raises KeyError
    codes = (idx[:, 0] + 3 * idx[:, 1] + 9 * idx[:, 2]).astype(np.uint8)
    return torch.as_tensor(codes, device=device)


def rotations_to_quat(rotations, rot_order):
    """Euler angles in DEGREES ``[n_frames, n_joints, 3]`` with the per-joint channel order ``rot_order``
    ``[n_joints, 3]`` of 'x' | 'y' | 'z' -> unrolled unit quaternions ``[n_frames, n_joints, 4]``
    (io/bvh.py:352-359)."""
    m = rt.Marshal(rotations)
    e = m.dev(rotations)
    if e.dim() != 3 or e.shape[-1] != 3:
        raise ValueError(f"rotations must have shape [n_frames, n_joints, 3], got {tuple(e.shape)}")
    e = e.contiguous()
    n_frames, n_joints = int(e.shape[0]), int(e.shape[1])
    codes = _order_codes_per_joint(rot_order, n_joints, m.device)
    out = m.new((n_frames, n_joints, 4))
    if n_frames > 0 and n_joints > 0:
        lib_bytes = n_frames * n_joints + ((n_frames + 127) // 128) * n_joints + 16  # pmb_unroll_workspace_bytes
        ws = torch.empty(lib_bytes, device=m.device, dtype=torch.uint8)
        rt.call("pmb_bvh_rotations_to_quat_f32", m.device, rt.ptr(e), rt.ptr(codes), n_frames, n_joints, rt.ptr(out),
                rt.ptr(ws), lib_bytes, m.stream())
    return m.out(out)


def get_data(data: dict):
    """Drop-in for ``BVH.get_data()`` given the reference's ``BVH.data`` dictionary (io/bvh.py:332-365):
    returns ``(rots, pos, parents, offsets, end_sites, end_sites_parents)`` with ``rots`` computed on the GPU."""
    rots = rotations_to_quat(data["rotations"], data["rot_order"])
    return rots, data["positions"], data["parents"], data["offsets"], data["end_sites"], data["end_sites_parents"]
