"""Drop-in for ``pymotion.ops.vector`` / ``vector_torch`` (/root/reference/pymotion/ops/vector.py:4-19)."""
from __future__ import annotations

from .. import _runtime as rt
from ..rotations.quat import _flat_count


def normalize(v, eps: float = 1e-8):
    """v / (|v| + eps) over the last axis."""
    m = rt.Marshal(v)
    t = m.dev(v).contiguous()
    if t.dim() < 1:
        raise ValueError("normalize needs at least one dimension")
    k = int(t.shape[-1])
    out = m.new(t.shape)
    n = _flat_count(t.shape[:-1])
    if n > 0 and k > 0:
        rt.call("pmb_vec_normalize_f32", m.device, rt.ptr(t), float(eps), rt.ptr(out), n, k, m.stream())
    return m.out(out)
