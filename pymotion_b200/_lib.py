"""ctypes binding of libpymotion_b200.so (the C ABI in include/pymotion_b200.h).

There is deliberately no fallback: if the library has not been built
(`python -m pymotion_b200._build`, or `__graft_entry__.build()`), importing any
op raises.  Nothing in this package computes on the CPU.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PMB_LIB_PATH") or os.path.join(_PKG, "libpymotion_b200.so")  # override: kernel experiments only

PMB_OK = 0
PMB_ERR_NULL, PMB_ERR_SHAPE, PMB_ERR_ALIGN, PMB_ERR_TOPOLOGY, PMB_ERR_CUDA, PMB_ERR_ROOT_OFFSET = -1, -2, -3, -4, -5, -6
PMB_MAX_JOINTS = 512

_vp, _i64, _i32, _f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.c_float

# name -> argtypes; every function returns int unless listed in _RESTYPES
SIGNATURES = {
    "pmb_version": [],
    "pmb_last_error": [],
    "pmb_last_variant": [],
    "pmb_status_string": [ctypes.c_int],
    "pmb_device_info": [_vp, _vp, _vp, ctypes.c_char_p, ctypes.c_int],
    "pmb_build_joint_program": [_vp, _i32, _vp],
    "pmb_build_track_schedule": [_vp, _i32, _i32, _i32, _vp, _i32, _vp],
    "pmb_reload_knobs": [],
    "pmb_fk_f32": [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp],
    "pmb_fk_quat_f32": [_vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp],
    "pmb_to_root_dual_quat_f32": [_vp, _vp, _i64, _vp, _vp, _vp, _i64, _i32, _vp, _vp],
    "pmb_from_root_dual_quat_f32": [_vp, _vp, _i64, _i32, _vp, _vp, _vp],
    "pmb_from_global_rotations_f32": [_vp, _vp, _i64, _i32, _vp, _vp],
    "pmb_fk_f32_host": [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64],
    "pmb_fk_quat_f32_host": [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp, _i64],
    "pmb_to_root_dual_quat_f32_host": [_vp, _vp, _vp, _vp, _i64, _i32, _vp, _i64],
    "pmb_from_root_dual_quat_f32_host": [_vp, _vp, _i64, _i32, _vp, _vp, _i64],
    "pmb_release_workspace": [],
    "pmb_quat_mul_f32": [_vp, _vp, _vp, _i64, _vp],
    "pmb_quat_mul_vec_f32": [_vp, _vp, _vp, _i64, _vp],
    "pmb_quat_length_f32": [_vp, _vp, _i64, _vp],
    "pmb_quat_normalize_f32": [_vp, _f32, _vp, _i64, _vp],
    "pmb_quat_conjugate_f32": [_vp, _vp, _i64, _vp],
    "pmb_quat_to_matrix_f32": [_vp, _vp, _i64, _vp],
    "pmb_quat_from_matrix_f32": [_vp, _vp, _i64, _vp],
    "pmb_dq_from_rotation_translation_f32": [_vp, _vp, _vp, _i64, _vp],
    "pmb_dq_from_translation_f32": [_vp, _vp, _i64, _vp],
    "pmb_dq_to_rotation_translation_f32": [_vp, _vp, _vp, _i64, _vp],
    "pmb_quat_from_angle_axis_f32": [_vp, _vp, _vp, _i64, _vp],
    "pmb_quat_from_scaled_angle_axis_f32": [_vp, _vp, _i64, _vp],
    "pmb_quat_from_euler_f32": [_vp, _vp, _i64, _vp, _i64, _vp],
    "pmb_quat_to_euler_f32": [_vp, _vp, _i64, _vp, _i64, _vp],
    "pmb_quat_to_angle_axis_f32": [_vp, _vp, _vp, _i64, _vp],
    "pmb_quat_to_scaled_angle_axis_f32": [_vp, _vp, _i64, _vp],
    "pmb_quat_slerp_f32": [_vp, _vp, _vp, _i64, _i32, _vp, _i64, _vp],
    "pmb_quat_from_to_f32": [_vp, _vp, _i32, _vp, _i64, _vp],
    "pmb_quat_from_to_axis_f32": [_vp, _vp, _vp, _i32, _vp, _i64, _vp],
    "pmb_unroll_workspace_bytes": [_i64, _i64],
    "pmb_unroll_f32": [_vp, _i32, _i64, _i64, _vp, _vp, _i64, _vp],
    "pmb_bvh_rotations_to_quat_f32": [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _vp],
    "pmb_dq_is_unit_f32": [_vp, _f32, _i64, _vp, _vp],
    "pmb_dq_normalize_f32": [_vp, _vp, _i64, _vp, _vp],
    "pmb_from_root_positions_f32": [_vp, _vp, _vp, _i64, _i32, _vp, _vp],
    "pmb_mirror_local_needs_scratch": [_vp, _i32],
    "pmb_mirror_local_f32": [_vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp, _vp],
    "pmb_mirror_to_local_f32": [_vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp],
    "pmb_vec_mirror_f32": [_vp, _i32, _vp, _i64, _vp],
    "pmb_root_center_f32": [_vp, _vp, _i64, _i32, _vp],
    "pmb_ortho6d_from_matrix_f32": [_vp, _vp, _i64, _vp],
    "pmb_ortho6d_from_quat_f32": [_vp, _vp, _i64, _vp],
    "pmb_ortho6d_to_matrix_f32": [_vp, _vp, _i64, _vp],
    "pmb_ortho6d_to_quat_f32": [_vp, _vp, _i64, _vp],
    "pmb_center_of_mass_f32": [_vp, _vp, _i64, _i64, _i32, _vp, _vp],
    "pmb_interpolate_positions_f32": [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _vp],
    "pmb_vec_normalize_f32": [_vp, _f32, _vp, _i64, _i32, _vp],
}
_RESTYPES = {"pmb_last_error": ctypes.c_char_p, "pmb_last_variant": ctypes.c_char_p, "pmb_status_string": ctypes.c_char_p, "pmb_release_workspace": None, "pmb_reload_knobs": None,
             "pmb_unroll_workspace_bytes": ctypes.c_int64}

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build the CUDA library with `python -m pymotion_b200._build` "
                "(pymotion_b200 has no CPU or eager fallback)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = header / library mismatch
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        _lib = lib
    return _lib


def last_error() -> str:
    return load().pmb_last_error().decode()


def check(status: int) -> None:
    """Map a pmb_status to the exception the reference would have raised."""
    if status == PMB_OK:
        return
    msg = last_error()
    if status == PMB_ERR_ROOT_OFFSET:
        raise AssertionError(msg)  # ops/skeleton.py:227 asserts
    if status in (PMB_ERR_NULL, PMB_ERR_SHAPE, PMB_ERR_ALIGN, PMB_ERR_TOPOLOGY):
        raise ValueError(msg)
    raise RuntimeError(msg)
