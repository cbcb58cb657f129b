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
