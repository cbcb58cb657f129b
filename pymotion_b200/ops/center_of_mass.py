"""Drop-in for ``pymotion.ops.center_of_mass`` / ``center_of_mass_torch``
(/root/reference/pymotion/ops/center_of_mass.py: human_center_of_mass :4, center_of_mass :52)."""
from __future__ import annotations

import torch

from .. import _runtime as rt
from ..rotations.quat import _flat_count


def center_of_mass(joints, weights):
    """sum_j joints[..., j, :] * weights[..., j]  ->  [..., 3]; weights [n_joints] or [..., n_joints]."""
    m = rt.Marshal(joints, weights)
    p = m.dev(joints)
    if p.dim() < 2 or p.shape[-1] != 3:
        raise ValueError(f"joints must have shape [..., n_joints, 3], got {tuple(p.shape)}")
    p = p.contiguous()
    lead, n_joints = tuple(p.shape[:-2]), int(p.shape[-2])
    w = m.dev(weights)
    if w.shape[-1] != n_joints:
        raise ValueError(f"weights must end in {n_joints} joints, got {tuple(w.shape)}")
    if w.dim() == 1:
        w, stride = w.contiguous(), 0
    else:
        w, stride = torch.broadcast_to(w, lead + (n_joints,)).contiguous(), n_joints
    out = m.new(lead + (3,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_center_of_mass_f32", m.device, rt.ptr(p), rt.ptr(w), stride, n, n_joints, rt.ptr(out), m.stream())
    return m.out(out)


def human_center_of_mass(joints_spine, joints_left_arm, joints_right_arm, joints_left_leg, joints_right_leg):
    """Spine 60 %, each arm 5 %, each leg 15 % of the body weight, spread evenly over the joints of each part
    (center_of_mass.py:4-49)."""
    m = rt.Marshal(joints_spine)
    parts = [m.dev(x) for x in (joints_spine, joints_left_arm, joints_right_arm, joints_left_leg, joints_right_leg)]
    shares = (0.6, 0.05, 0.05, 0.15, 0.15)
    weights = [share / part.shape[-2] for part, share in zip(parts, shares) for _ in range(part.shape[-2])]
    joints = torch.cat(parts, dim=-2)
    out = center_of_mass(joints, torch.tensor(weights, dtype=torch.float32, device=m.device))
    return m.out(out)
