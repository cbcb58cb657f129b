"""Dual-quaternion primitives, layout [..., (w_r x_r y_r z_r | w_d x_d y_d z_d)]:
drop-in for ``pymotion.rotations.dual_quat`` / ``dual_quat_torch``
(/root/reference/pymotion/rotations/dual_quat.py: from_rotation_translation :12,
from_translation :39, to_rotation_translation :62).  One CUDA kernel per call."""
from __future__ import annotations

import torch

from .. import _runtime as rt
from .quat import _flat_count, _pair


def from_rotation_translation(rotations, translations):
    """q_r = rotations, q_d = 0.5 * ((0, t) (x) q_r)  ->  [..., 8] (dual_quat.py:12-36)."""
    m = rt.Marshal(rotations, translations)
    r, t, lead = _pair(m, rotations, translations, 4, 3)
    dq = m.new(lead + (8,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_dq_from_rotation_translation_f32", m.device, rt.ptr(r), rt.ptr(t), rt.ptr(dq), n, m.stream())
    return m.out(dq)


def from_translation(translations):
    """Identity rotation + translation -> [..., 8] (dual_quat.py:39-59)."""
    m = rt.Marshal(translations)
    t = m.dev(translations)
    if t.shape[-1] != 3:
        raise ValueError(f"expected [..., 3], got {tuple(t.shape)}")
    t = t.contiguous()
    dq = m.new(tuple(t.shape[:-1]) + (8,))
    n = _flat_count(t.shape[:-1])
    if n > 0:
        rt.call("pmb_dq_from_translation_f32", m.device, rt.ptr(t), rt.ptr(dq), n, m.stream())
    return m.out(dq)


def to_rotation_translation(dq):
    """[..., 8] -> (rotations [..., 4], translations [..., 3]) (dual_quat.py:62-83)."""
    m = rt.Marshal(dq)
    d = m.dev(dq)
    if d.shape[-1] != 8:
        raise ValueError(f"expected [..., 8], got {tuple(d.shape)}")
    d = d.contiguous()
    lead = tuple(d.shape[:-1])
    r, t = m.new(lead + (4,)), m.new(lead + (3,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_dq_to_rotation_translation_f32", m.device, rt.ptr(d), rt.ptr(r), rt.ptr(t), n, m.stream())
    return m.out(r), m.out(t)
