"""6-D rotation representation, layout [..., 3, 2] = the first two columns of the rotation matrix: drop-in
for ``pymotion.rotations.ortho6d`` / ``ortho6d_torch`` (/root/reference/pymotion/rotations/ortho6d.py:
from_quat :14, from_matrix :31, to_quat :50, to_matrix :67).  One CUDA kernel per call."""
from __future__ import annotations

from .quat import _unary


def from_quat(quaternions):
    """[..., 4] -> [..., 3, 2] (ortho6d.py:14-28)."""
    return _unary("pmb_ortho6d_from_quat_f32", quaternions, 4, (3, 2))


def from_matrix(rotmats):
    """[..., 3, 3] -> [..., 3, 2] (ortho6d.py:31-47; a contiguous copy, the NumPy reference returns a view)."""
    return _unary("pmb_ortho6d_from_matrix_f32", rotmats, (3, 3), (3, 2))


def to_quat(ortho6D):
    """[..., 3, 2] -> [..., 4] (ortho6d.py:50-64)."""
    return _unary("pmb_ortho6d_to_quat_f32", ortho6D, (3, 2), (4,))


def to_matrix(ortho6D):
    """[..., 3, 2] -> [..., 3, 3] by Gram-Schmidt on the two columns (ortho6d.py:67-90)."""
    return _unary("pmb_ortho6d_to_matrix_f32", ortho6D, (3, 2), (3, 3))
