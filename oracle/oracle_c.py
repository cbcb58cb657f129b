"""ctypes front-end of oracle/oracle_c.c (TEST INFRASTRUCTURE ONLY).

Fast, multi-threaded restatement of the reference path used to check
full-size batches and as the all-cores CPU baseline of bench.py."""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.LIB
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(build_oracle.SRC):
            path = build_oracle.build()
        _lib = ctypes.CDLL(path)
        i64, i32, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_void_p
        _lib.orc_fk_f32.argtypes = [vp, vp, i64, vp, i64, vp, i64, i32, vp, vp]
        _lib.orc_to_root_dual_quat_f32.argtypes = [vp, vp, i64, vp, vp, i64, i32, vp]
        _lib.orc_from_root_dual_quat_f64.argtypes = [vp, vp, i64, i32, vp, vp]
        for f in (_lib.orc_fk_f32, _lib.orc_to_root_dual_quat_f32, _lib.orc_from_root_dual_quat_f64):
            f.restype = ctypes.c_int
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _flat(rot, gpos):
    rot = np.ascontiguousarray(rot, dtype=np.float32)
    lead, n_joints = rot.shape[:-2], rot.shape[-2]
    n_frames = int(np.prod(lead)) if lead else 1
    gpos = np.ascontiguousarray(np.broadcast_to(np.asarray(gpos, dtype=np.float32), lead + (3,)))
    return rot, gpos, lead, n_frames, n_joints


def fk(rot, global_pos, offsets, parents):
    rot, gpos, lead, n_frames, n_joints = _flat(rot, global_pos)
    offsets = np.ascontiguousarray(offsets, dtype=np.float32)
    off_stride = 0 if offsets.ndim == 2 else n_joints * 3
    if offsets.ndim != 2:
        offsets = np.ascontiguousarray(np.broadcast_to(offsets, lead + (n_joints, 3)))
    par = np.ascontiguousarray(parents, dtype=np.int64)
    pos = np.empty(lead + (n_joints, 3))
    rotm = np.empty(lead + (n_joints, 3, 3))
    rc = lib().orc_fk_f32(_p(rot), _p(gpos), 3, _p(offsets), off_stride, _p(par), n_frames, n_joints, _p(pos), _p(rotm))
    if rc:
        raise ValueError(f"orc_fk_f32 failed: {rc}")
    return pos, rotm


def to_root_dual_quat(rotations, global_pos, parents, offsets):
    rot, gpos, lead, n_frames, n_joints = _flat(rotations, global_pos)
    offsets = np.ascontiguousarray(offsets, dtype=np.float32)
    par = np.ascontiguousarray(parents, dtype=np.int64)
    dq = np.empty(lead + (n_joints, 8))
    rc = lib().orc_to_root_dual_quat_f32(_p(rot), _p(gpos), 3, _p(par), _p(offsets), n_frames, n_joints, _p(dq))
    if rc == -3:
        raise AssertionError("offsets[0] must be zero")
    if rc:
        raise ValueError(f"orc_to_root_dual_quat_f32 failed: {rc}")
    return dq


def from_root_dual_quat(dq, parents):
    dq = np.ascontiguousarray(dq, dtype=np.float64)
    lead, n_joints = dq.shape[:-2], dq.shape[-2]
    n_frames = int(np.prod(lead)) if lead else 1
    par = np.ascontiguousarray(parents, dtype=np.int64)
    trans = np.empty(lead + (n_joints, 3))
    rots = np.empty(lead + (n_joints, 4))
    rc = lib().orc_from_root_dual_quat_f64(_p(dq), _p(par), n_frames, n_joints, _p(trans), _p(rots))
    if rc:
        raise ValueError(f"orc_from_root_dual_quat_f64 failed: {rc}")
    return trans, rots
