"""Quaternion primitives, layout [..., (w, x, y, z)]: drop-in for the functions of
``pymotion.rotations.quat`` / ``quat_torch`` that sit on the fk / dual-quaternion
path (/root/reference/pymotion/rotations/quat.py: mul :337, mul_vec :320,
length :364, inverse :379, conjugate :396, normalize :411, to_matrix :276,
from_matrix :85).  One CUDA kernel per call through the C ABI; operands are
broadcast against each other like the NumPy originals."""
from __future__ import annotations

import math

import torch

from .. import _runtime as rt


def _flat_count(shape) -> int:
    return int(math.prod(shape)) if len(shape) else 1


def _pair(m: rt.Marshal, a, b, tail_a: int, tail_b: int):
    ta, tb = m.dev(a), m.dev(b)
    if ta.shape[-1] != tail_a or tb.shape[-1] != tail_b:
        raise ValueError(f"expected last dims {tail_a} and {tail_b}, got {tuple(ta.shape)} and {tuple(tb.shape)}")
    lead = torch.broadcast_shapes(ta.shape[:-1], tb.shape[:-1])
    ta = torch.broadcast_to(ta, lead + (tail_a,)).contiguous()
    tb = torch.broadcast_to(tb, lead + (tail_b,)).contiguous()
    return ta, tb, tuple(lead)


def _unary(name: str, x, tail_in: int, tail_out, *extra):
    m = rt.Marshal(x)
    t = m.dev(x)
    tin = (tail_in,) if isinstance(tail_in, int) else tuple(tail_in)
    if tuple(t.shape[-len(tin):]) != tin:
        raise ValueError(f"expected trailing shape {tin}, got {tuple(t.shape)}")
    t = t.contiguous()
    lead = tuple(t.shape[: t.dim() - len(tin)])
    out = m.new(lead + tuple(tail_out))
    n = _flat_count(lead)
    if n > 0:
        rt.call(name, m.device, rt.ptr(t), *extra, rt.ptr(out), n, m.stream())
    return m.out(out)


def mul(q0, q1):
    """Hamilton product q0 (x) q1 (quat.py:337-361)."""
    m = rt.Marshal(q0, q1)
    a, b, lead = _pair(m, q0, q1, 4, 4)
    out = m.new(lead + (4,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_quat_mul_f32", m.device, rt.ptr(a), rt.ptr(b), rt.ptr(out), n, m.stream())
    return m.out(out)


def mul_vec(q, v):
    """Rotate vectors v [..., 3] by quaternions q [..., 4] (quat.py:320-334)."""
    m = rt.Marshal(q, v)
    a, b, lead = _pair(m, q, v, 4, 3)
    out = m.new(lead + (3,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_quat_mul_vec_f32", m.device, rt.ptr(a), rt.ptr(b), rt.ptr(out), n, m.stream())
    return m.out(out)


def length(quaternions):
    """Euclidean norm, shape [...] (quat.py:364-376)."""
    return _unary("pmb_quat_length_f32", quaternions, 4, ())


def normalize(quaternions, eps: float = 1e-8):
    """q / (|q| + eps) (quat.py:411-423)."""
    m = rt.Marshal(quaternions)
    t = m.dev(quaternions)
    if t.shape[-1] != 4:
        raise ValueError(f"expected [..., 4], got {tuple(t.shape)}")
    t = t.contiguous()
    out = m.new(t.shape)
    n = _flat_count(t.shape[:-1])
    if n > 0:
        rt.call("pmb_quat_normalize_f32", m.device, rt.ptr(t), float(eps), rt.ptr(out), n, m.stream())
    return m.out(out)


def conjugate(quaternions):
    """(w, -x, -y, -z) (quat.py:396-408)."""
    return _unary("pmb_quat_conjugate_f32", quaternions, 4, (4,))


def inverse(quaternions):
    """Inverse of a UNIT quaternion = its conjugate (quat.py:379-393)."""
    return conjugate(quaternions)


def to_matrix(quaternions):
    """[..., 4] -> [..., 3, 3], row-major, no normalisation inside (quat.py:276-317)."""
    return _unary("pmb_quat_to_matrix_f32", quaternions, 4, (3, 3))


def from_matrix(rotmats):
    """[..., 3, 3] -> [..., 4], four-branch extraction + normalize (quat.py:85-156)."""
    return _unary("pmb_quat_from_matrix_f32", rotmats, (3, 3), (4,))


# ---------------------------------------------------------------------------------------------
# The rest of pymotion.rotations.quat (off the fk hot path, same drop-in rules)
# ---------------------------------------------------------------------------------------------
_AXES = {"x": 0, "y": 1, "z": 2}


def _order_codes(order, lead, device):
    """'x'|'y'|'z' triples -> one byte per element (o0 + 3*o1 + 9*o2) on the device; a single shared order
    (all rows equal) travels as one byte with stride 0.  Host-side text handling only: no arithmetic."""
    import numpy as np

    arr = np.asarray(order)
    if arr.shape[-1:] != (3,):
        raise ValueError(f"order must have shape [..., 3], got {arr.shape}")
    if tuple(arr.shape[:-1]) != tuple(lead):
        # the reference asserts this (quat.py:60-62, :177-179)
        raise AssertionError("euler / quaternions and order must have the same shape except for the last dimension")
    idx = np.full(arr.shape, -1, dtype=np.int64)
    for ch, k in _AXES.items():
        idx[arr == ch] = k
    if (idx < 0).any():
        raise KeyError("order entries must be 'x', 'y' or 'z'")  # the reference's dict lookup raises KeyError
    codes = (idx[..., 0] + 3 * idx[..., 1] + 9 * idx[..., 2]).astype(np.uint8).reshape(-1)
    if codes.size and (codes == codes[0]).all():
        return torch.as_tensor(codes[:1], device=device), 0
    return torch.as_tensor(codes, device=device), 1


def from_angle_axis(angle, axis):
    """(cos(a/2), sin(a/2) * axis); angle [..., 1], axis [..., 3] unit (quat.py:24-40)."""
    m = rt.Marshal(angle, axis)
    a, b, lead = _pair(m, angle, axis, 1, 3)
    out = m.new(lead + (4,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_quat_from_angle_axis_f32", m.device, rt.ptr(a), rt.ptr(b), rt.ptr(out), n, m.stream())
    return m.out(out)


def from_scaled_angle_axis(scaledaxis):
    """Axis scaled by the angle -> quaternion (quat.py:6-21); the null vector gives nan like the reference."""
    return _unary("pmb_quat_from_scaled_angle_axis_f32", scaledaxis, 3, (4,))


def from_euler(euler, order):
    """Intrinsic Euler angles [..., 3] (radians) with per-element axis order [..., 3] of 'x'|'y'|'z'
    (quat.py:43-82)."""
    m = rt.Marshal(euler)
    e = m.dev(euler)
    if e.shape[-1] != 3:
        raise ValueError(f"expected [..., 3], got {tuple(e.shape)}")
    e = e.contiguous()
    lead = tuple(e.shape[:-1])
    codes, stride = _order_codes(order, lead, m.device)
    out = m.new(lead + (4,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_quat_from_euler_f32", m.device, rt.ptr(e), rt.ptr(codes), stride, rt.ptr(out), n, m.stream())
    return m.out(out)


def to_euler(quaternions, order):
    """Quaternion -> intrinsic Euler angles in [0, 2 pi) for the given order; no gimbal handling (quat.py:159-227)."""
    m = rt.Marshal(quaternions)
    q = m.dev(quaternions)
    if q.shape[-1] != 4:
        raise ValueError(f"expected [..., 4], got {tuple(q.shape)}")
    q = q.contiguous()
    lead = tuple(q.shape[:-1])
    codes, stride = _order_codes(order, lead, m.device)
    out = m.new(lead + (3,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_quat_to_euler_f32", m.device, rt.ptr(q), rt.ptr(codes), stride, rt.ptr(out), n, m.stream())
    return m.out(out)


def to_angle_axis(quaternions):
    """-> (angle [..., 1], axis [..., 3]); zero axis where sin(angle/2) <= 1e-8 (quat.py:247-273)."""
    m = rt.Marshal(quaternions)
    q = m.dev(quaternions)
    if q.shape[-1] != 4:
        raise ValueError(f"expected [..., 4], got {tuple(q.shape)}")
    q = q.contiguous()
    lead = tuple(q.shape[:-1])
    angle, axis = m.new(lead + (1,)), m.new(lead + (3,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_quat_to_angle_axis_f32", m.device, rt.ptr(q), rt.ptr(angle), rt.ptr(axis), n, m.stream())
    return m.out(angle), m.out(axis)


def to_scaled_angle_axis(quaternions):
    """angle * axis (quat.py:230-244)."""
    return _unary("pmb_quat_to_scaled_angle_axis_f32", quaternions, 4, (3,))


def _unroll(x, axis, width: int):
    m = rt.Marshal(x)
    t = m.dev(x)
    if t.shape[-1] != width:
        raise ValueError(f"expected [..., {width}], got {tuple(t.shape)}")
    nd = t.dim()
    axis = axis + nd if axis < 0 else axis
    if not 0 <= axis < nd - 1:
        raise ValueError(f"unroll axis {axis} out of range for shape {tuple(t.shape)}")
    moved = torch.movedim(t, axis, 0).contiguous()
    n_steps = moved.shape[0]
    n_cols = _flat_count(moved.shape[1:-1])
    out = m.new(moved.shape)
    if n_steps > 0 and n_cols > 0:
        lib_bytes = int(rt._lib.load().pmb_unroll_workspace_bytes(n_steps, n_cols))
        work = torch.empty(lib_bytes, device=m.device, dtype=torch.uint8)
        rt.call("pmb_unroll_f32", m.device, rt.ptr(moved), width, n_steps, n_cols, rt.ptr(out), rt.ptr(work),
                lib_bytes, m.stream())
    return m.out(torch.movedim(out, 0, axis).contiguous())


def unroll(quaternions, axis=None, dim=None):
    """Remove sign flips along `axis` (torch twin: `dim`): entry i is negated iff its dot product with the
    already unrolled entry i-1 is negative (quat.py:426-462).  Returns a new array -- the NumPy reference
    flips its input in place through a view."""
    if axis is None:
        axis = dim
    if axis is None:
        raise TypeError("unroll() missing the axis / dim argument")
    return _unroll(quaternions, int(axis), 4)


def slerp(q0, q1, t, shortest: bool = True):
    """Spherical linear interpolation; t a float or an array [..., 1] (quat.py:465-501)."""
    m = rt.Marshal(q0, q1)
    a, b, lead = _pair(m, q0, q1, 4, 4)
    n = _flat_count(lead)
    if isinstance(t, (int, float)):
        tt, stride = torch.full((1,), float(t), device=m.device, dtype=torch.float32), 0
    else:
        tt = m.dev(t)
        if tt.dim() == 0:
            tt, stride = tt.reshape(1).contiguous(), 0
        else:
            if tt.shape[-1] != 1:
                raise ValueError(f"t must be a float or have shape [..., 1], got {tuple(tt.shape)}")
            lead_t = torch.broadcast_shapes(lead, tt.shape[:-1])
            if tuple(lead_t) != tuple(lead):
                a = torch.broadcast_to(a, lead_t + (4,)).contiguous()
                b = torch.broadcast_to(b, lead_t + (4,)).contiguous()
                lead, n = tuple(lead_t), _flat_count(lead_t)
            tt, stride = torch.broadcast_to(tt, lead + (1,)).contiguous(), 1
    out = m.new(lead + (4,))
    if n > 0:
        rt.call("pmb_quat_slerp_f32", m.device, rt.ptr(a), rt.ptr(b), rt.ptr(tt), stride, int(bool(shortest)),
                rt.ptr(out), n, m.stream())
    return m.out(out)


def _same_shape3(*arrays):
    shapes = [tuple(a.shape) for a in arrays]
    if any(s[-1:] != (3,) for s in shapes):
        raise AssertionError("Input vectors must have shape [..., 3]")  # quat.py:521, :601
    if any(s != shapes[0] for s in shapes):
        raise AssertionError("Input vectors must have the same shape")  # quat.py:522, :602-603
    return shapes[0][:-1]


def from_to(v1, v2, normalize_input: bool = True):
    """Rotation taking direction v1 onto v2; identity for parallel, a half turn for anti-parallel inputs
    (quat.py:504-576)."""
    m = rt.Marshal(v1, v2)
    a, b = m.dev(v1).contiguous(), m.dev(v2).contiguous()
    lead = _same_shape3(a, b)
    out = m.new(tuple(lead) + (4,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_quat_from_to_f32", m.device, rt.ptr(a), rt.ptr(b), int(bool(normalize_input)), rt.ptr(out), n,
                m.stream())
    return m.out(out)


def from_to_axis(v1, v2, rot_axis, normalize_input: bool = True):
    """Same, about a GIVEN rotation axis (quat.py:579-650)."""
    m = rt.Marshal(v1, v2, rot_axis)
    a, b, c = m.dev(v1).contiguous(), m.dev(v2).contiguous(), m.dev(rot_axis).contiguous()
    lead = _same_shape3(a, b, c)
    out = m.new(tuple(lead) + (4,))
    n = _flat_count(lead)
    if n > 0:
        rt.call("pmb_quat_from_to_axis_f32", m.device, rt.ptr(a), rt.ptr(b), rt.ptr(c), int(bool(normalize_input)),
                rt.ptr(out), n, m.stream())
    return m.out(out)
