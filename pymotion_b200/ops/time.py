"""Drop-in for ``pymotion.ops.time`` / ``time_torch`` (/root/reference/pymotion/ops/time.py:4-66)."""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _runtime as rt


def interpolate_positions(sample_times, original_times, positions, axis: int, method: str = "linear"):
    """Linear interpolation of ``positions [..., 3]`` along ``axis`` at ``sample_times``; the interval of a
    sample comes from ``np.searchsorted`` (left) clamped to the first / last interval, so samples outside the
    original range extrapolate (time.py:49-64).  Times are handled in float64.  Works for any ``axis`` (the
    reference's weight broadcasting only lines up when ``axis`` is the second-to-last axis or the array is
    [T, 3])."""
    assert method == "linear", "Only linear interpolation is supported yet."
    m = rt.Marshal(positions)
    p = m.dev(positions)
    nd = p.dim()
    axis = axis + nd if axis < 0 else axis
    if not 0 <= axis < nd:
        raise ValueError(f"axis {axis} out of range for positions of shape {tuple(p.shape)}")

    def times(x):
        if isinstance(x, torch.Tensor):
            return x.detach().to(device=m.device, dtype=torch.float64).contiguous().reshape(-1)
        return torch.as_tensor(np.asarray(x, dtype=np.float64), device=m.device).reshape(-1)

    ts, t0 = times(sample_times), times(original_times)
    assert p.shape[axis] == t0.shape[0], (
        "Wrong shape of data. Positions along the axis dimension must be equal to the length of original_times.")
    p = p.contiguous()
    outer = int(math.prod(p.shape[:axis])) if axis > 0 else 1
    inner = int(math.prod(p.shape[axis + 1:])) if axis + 1 < nd else 1
    n_orig, n_samp = int(t0.shape[0]), int(ts.shape[0])
    out = m.new(tuple(p.shape[:axis]) + (n_samp,) + tuple(p.shape[axis + 1:]))
    if out.numel() > 0:
        idx = torch.empty(n_samp, device=m.device, dtype=torch.int32)
        wts = torch.empty(n_samp, device=m.device, dtype=torch.float32)
        rt.call("pmb_interpolate_positions_f32", m.device, rt.ptr(ts), rt.ptr(t0), rt.ptr(p), outer, n_orig, n_samp,
                inner, rt.ptr(out), rt.ptr(idx), rt.ptr(wts), m.stream())
    return m.out(out)
