"""Skeleton ops: drop-in for ``pymotion.ops.skeleton`` / ``skeleton_torch`` on the
path BASELINE.json names -- ``fk``, ``to_root_dual_quat``, ``from_root_dual_quat``
(plus ``from_global_rotations``, the step that follows ``fk`` in every in-repo
caller of the reference).

Same names, positional order and return order as the reference
(/root/reference/pymotion/ops/skeleton.py:16, :207, :173, :64).  Each call is one
CUDA kernel launched through the C ABI (include/pymotion_b200.h) on the current
stream of the tensors' device.  Deliberate, documented differences:

* results are contiguous float32 tensors (input dtype restored on return), not
  float64 views of a 4x4 buffer;
* the joint count is ``shape[-2]`` (the NumPy reference reads ``shape[1]`` in the
  dual-quaternion pair, ops/skeleton.py:188/:228, which is only right for 3-D input);
* ``parents[i]`` must be in ``[0, i)`` for ``i >= 1`` -> ``ValueError`` otherwise.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _runtime as rt


def _lead_frames(shape_lead) -> int:
    return int(math.prod(shape_lead)) if len(shape_lead) else 1


def _broadcast_rows(m: rt.Marshal, x, lead, tail) -> tuple[torch.Tensor, int]:
    """Return (tensor, frame_stride).  A single row shared by every frame keeps
    stride 0 (no materialised broadcast); anything else is expanded to
    lead + tail and made contiguous."""
    t = m.dev(x)
    n_tail = int(math.prod(tail))
    if t.numel() == n_tail and tuple(t.shape[-len(tail):]) == tuple(tail):
        return t.reshape(tail).contiguous(), 0
    t = torch.broadcast_to(t, tuple(lead) + tuple(tail)).contiguous()
    return t, n_tail


def fk(rot, global_pos, offsets, parents):
    """Forward kinematics (reference: ops/skeleton.py:16-61).

    Parameters
    ----------
    rot : [..., n_joints, 4]   local rotations (w, x, y, z); normalised inside like the reference
    global_pos : [..., 3]      root position (broadcast against the leading dims of ``rot``)
    offsets : [n_joints, 3] or [..., n_joints, 3]
    parents : [n_joints] ints  ``parents[0]`` is ignored

    Returns
    -------
    positions : [..., n_joints, 3]
    rotmats : [..., n_joints, 3, 3]   row-major, ``v' = M v``
    """
    m = rt.Marshal(rot, global_pos, offsets)
    q = m.dev(rot)
    if q.dim() < 2 or q.shape[-1] != 4:
        raise ValueError(f"rot must have shape [..., n_joints, 4], got {tuple(q.shape)}")
    q = q.contiguous()
    lead, n_joints = tuple(q.shape[:-2]), int(q.shape[-2])
    par = rt.host_parents(parents)
    if par.shape[0] != n_joints:
        raise ValueError(f"parents has {par.shape[0]} entries but rot has {n_joints} joints")
    n_frames = _lead_frames(lead)
    gp, g_stride = _broadcast_rows(m, global_pos, lead, (3,))
    off, o_stride = _broadcast_rows(m, offsets, lead, (n_joints, 3))
    pos = m.new(lead + (n_joints, 3))
    rotm = m.new(lead + (n_joints, 3, 3))
    if n_frames > 0:
        rt.call("pmb_fk_f32", m.device, rt.ptr(q), rt.ptr(gp), g_stride, rt.ptr(off), o_stride,
                par.ctypes.data, n_frames, n_joints, rt.ptr(pos), rt.ptr(rotm), m.stream())
    return m.out(pos), m.out(rotm)


def fk_quat(rot, global_pos, offsets, parents):
    """``fk`` that returns global QUATERNIONS instead of rotation matrices:
    ``(positions, global_rots)`` with ``global_rots == quat.from_matrix(rotmats)``
    (sign included: it goes through the same branch selection, quat.py:85-156).
    This is what ``mirror`` / ``from_root_positions`` of the reference compute right
    after ``fk`` (ops/skeleton.py:322-323, :134-140); 44 J instead of 64 J bytes per pose."""
    m = rt.Marshal(rot, global_pos, offsets)
    q = m.dev(rot)
    if q.dim() < 2 or q.shape[-1] != 4:
        raise ValueError(f"rot must have shape [..., n_joints, 4], got {tuple(q.shape)}")
    q = q.contiguous()
    lead, n_joints = tuple(q.shape[:-2]), int(q.shape[-2])
    par = rt.host_parents(parents)
    if par.shape[0] != n_joints:
        raise ValueError(f"parents has {par.shape[0]} entries but rot has {n_joints} joints")
    n_frames = _lead_frames(lead)
    gp, g_stride = _broadcast_rows(m, global_pos, lead, (3,))
    off, o_stride = _broadcast_rows(m, offsets, lead, (n_joints, 3))
    pos = m.new(lead + (n_joints, 3))
    grot = m.new(lead + (n_joints, 4))
    if n_frames > 0:
        rt.call("pmb_fk_quat_f32", m.device, rt.ptr(q), rt.ptr(gp), g_stride, rt.ptr(off), o_stride,
                par.ctypes.data, n_frames, n_joints, rt.ptr(pos), rt.ptr(grot), m.stream())
    return m.out(pos), m.out(grot)


def to_root_dual_quat(rotations, global_pos, parents, offsets):
    """Root-centred dual quaternions (reference: ops/skeleton.py:207-244).

    NOTE the argument order (parents before offsets), as in the reference.
    Raises ``AssertionError`` if ``offsets[0] != 0`` (ops/skeleton.py:227).
    Returns ``dq`` of shape [..., n_joints, 8].
    """
    m = rt.Marshal(rotations, global_pos, offsets)
    q = m.dev(rotations)
    if q.dim() < 2 or q.shape[-1] != 4:
        raise ValueError(f"rotations must have shape [..., n_joints, 4], got {tuple(q.shape)}")
    q = q.contiguous()
    lead, n_joints = tuple(q.shape[:-2]), int(q.shape[-2])
    par = rt.host_parents(parents)
    if par.shape[0] != n_joints:
        raise ValueError(f"parents has {par.shape[0]} entries but rotations has {n_joints} joints")
    off = m.dev(offsets)
    if tuple(off.shape) != (n_joints, 3):
        # the reference's assert compares offsets[0] with zeros(3) and fails for per-frame offsets too
        raise AssertionError(f"offsets must have shape [{n_joints}, 3] with offsets[0] == 0, got {tuple(off.shape)}")
    off = off.contiguous()
    # the reference asserts on the VALUE of offsets[0]; reading 12 bytes back is the only sync of this call
    if isinstance(offsets, torch.Tensor):
        off0 = offsets[0].detach().to("cpu", torch.float32).contiguous().numpy()
    else:
        off0 = np.ascontiguousarray(np.asarray(offsets)[0], dtype=np.float32)
    n_frames = _lead_frames(lead)
    gp, g_stride = _broadcast_rows(m, global_pos, lead, (3,))
    dq = m.new(lead + (n_joints, 8))
    if n_frames > 0:
        rt.call("pmb_to_root_dual_quat_f32", m.device, rt.ptr(q), rt.ptr(gp), g_stride, par.ctypes.data,
                rt.ptr(off), off0.ctypes.data, n_frames, n_joints, rt.ptr(dq), m.stream())
    return m.out(dq)


def from_root_dual_quat(dq, parents):
    """Inverse of :func:`to_root_dual_quat` (reference: ops/skeleton.py:173-204).

    Returns ``(translations [..., n_joints, 3], rotations [..., n_joints, 4])`` --
    in that order, which is what the reference returns (:204) even though its
    docstring lists them the other way round.
    """
    m = rt.Marshal(dq)
    d = m.dev(dq)
    if d.dim() < 2 or d.shape[-1] != 8:
        raise ValueError(f"dq must have shape [..., n_joints, 8], got {tuple(d.shape)}")
    d = d.contiguous()
    lead, n_joints = tuple(d.shape[:-2]), int(d.shape[-2])
    par = rt.host_parents(parents)
    if par.shape[0] != n_joints:
        raise ValueError(f"parents has {par.shape[0]} entries but dq has {n_joints} joints")
    trans = m.new(lead + (n_joints, 3))
    rots = m.new(lead + (n_joints, 4))
    if d.numel() > 0:
        rt.call("pmb_from_root_dual_quat_f32", m.device, rt.ptr(d), par.ctypes.data, _lead_frames(lead), n_joints,
                rt.ptr(trans), rt.ptr(rots), m.stream())
    return m.out(trans), m.out(rots)


def from_global_rotations(global_quats, parents):
    """Global -> local rotations (reference: ops/skeleton.py:64-93):
    ``local_i = conj(global_parents[i]) (x) global_i``, root unchanged."""
    m = rt.Marshal(global_quats)
    g = m.dev(global_quats)
    if g.dim() < 2 or g.shape[-1] != 4:
        raise ValueError(f"global_quats must have shape [..., n_joints, 4], got {tuple(g.shape)}")
    g = g.contiguous()
    lead, n_joints = tuple(g.shape[:-2]), int(g.shape[-2])
    par = rt.host_parents(parents)
    if par.shape[0] != n_joints:
        raise ValueError(f"parents has {par.shape[0]} entries but global_quats has {n_joints} joints")
    out = m.new(lead + (n_joints, 4))
    if g.numel() > 0:
        rt.call("pmb_from_global_rotations_f32", m.device, rt.ptr(g), par.ctypes.data, _lead_frames(lead), n_joints,
                rt.ptr(out), m.stream())
    return m.out(out)


def fk_host(rot, global_pos, offsets, parents, out=None, chunk_frames: int = 0):
    """End-to-end ``fk`` on HOST buffers (the reference's calling convention: arrays
    live in host memory).  ``rot`` [F, J, 4] and ``global_pos`` [F, 3] are float32
    NumPy arrays or CPU tensors -- page-locked for full PCIe rate --, ``offsets`` is
    [J, 3].  The frame axis is cut into chunks that are copied in, computed and
    copied out on two streams so H2D, kernel and D2H overlap (pmb_fk_f32_host).

    ``out=(positions, rotmats)`` may supply (pinned) CPU tensors to fill; otherwise
    pinned tensors are allocated.  Returns CPU tensors (NumPy arrays if ``rot`` was one).
    """
    from .. import _lib

    as_numpy = isinstance(rot, np.ndarray)
    q = torch.as_tensor(rot)
    gp = torch.as_tensor(global_pos)
    off = torch.as_tensor(offsets)
    if q.is_cuda or gp.is_cuda:
        raise ValueError("fk_host takes host buffers; use fk() for device tensors")
    if q.dtype != torch.float32 or gp.dtype != torch.float32:
        raise TypeError("fk_host needs float32 host buffers")
    if q.dim() != 3 or q.shape[-1] != 4:
        raise ValueError(f"rot must have shape [n_frames, n_joints, 4], got {tuple(q.shape)}")
    n_frames, n_joints = int(q.shape[0]), int(q.shape[1])
    if tuple(gp.shape) != (n_frames, 3) or tuple(off.shape) != (n_joints, 3):
        raise ValueError("global_pos must be [n_frames, 3] and offsets [n_joints, 3]")
    par = rt.host_parents(parents)
    if par.shape[0] != n_joints:
        raise ValueError(f"parents has {par.shape[0]} entries but rot has {n_joints} joints")
    q, gp = q.contiguous(), gp.contiguous()
    off = off.to(torch.float32).contiguous()
    if out is None:
        pin = torch.cuda.is_available()
        pos = torch.empty((n_frames, n_joints, 3), dtype=torch.float32, pin_memory=pin)
        rotm = torch.empty((n_frames, n_joints, 3, 3), dtype=torch.float32, pin_memory=pin)
    else:
        pos, rotm = out
        if tuple(pos.shape) != (n_frames, n_joints, 3) or tuple(rotm.shape) != (n_frames, n_joints, 3, 3):
            raise ValueError("out buffers have the wrong shape")
        if not (pos.is_contiguous() and rotm.is_contiguous() and pos.dtype == rotm.dtype == torch.float32):
            raise ValueError("out buffers must be contiguous float32 CPU tensors")
    device = rt.default_device()
    if n_frames > 0:
        with torch.cuda.device(device):
            _lib.check(_lib.load().pmb_fk_f32_host(q.data_ptr(), gp.data_ptr(), off.data_ptr(), par.ctypes.data,
                                                   n_frames, n_joints, pos.data_ptr(), rotm.data_ptr(),
                                                   int(chunk_frames)))
    if as_numpy and out is None:
        return pos.numpy(), rotm.numpy()
    return pos, rotm
