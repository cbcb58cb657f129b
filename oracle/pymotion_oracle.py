"""CPU oracle for the pymotion FK / root dual-quaternion hot path.

TEST INFRASTRUCTURE ONLY.  This file is a NumPy restatement of the reference
algorithm (UPC-ViRVIG/pymotion v0.2.3) for the one path this repository
accelerates.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``pymotion_b200`` never does (and raises if its CUDA library is
missing rather than falling back to anything in here).

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the real reference
from ``/root/reference`` and writes its outputs to ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those
fixtures and against the hand-written golden arrays of the reference's own
tests (``pymotion/ops/tests/test_skeleton.py``, ``pymotion/rotations/tests``).

Every function cites the reference lines it restates.  dtype behaviour is part
of the contract and is reproduced on purpose (e.g. ``fk`` composes in float64
from input-precision local matrices because the reference allocates its 4x4
buffer with the NumPy default dtype).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "quat_length", "quat_normalize", "quat_to_matrix", "quat_from_matrix",
    "quat_mul", "quat_mul_vec", "quat_conjugate", "quat_inverse",
    "dq_from_rotation_translation", "dq_to_rotation_translation",
    "dq_from_translation", "fk", "to_root_dual_quat", "from_root_dual_quat",
    "from_global_rotations",
]


# ----------------------------------------------------------------------------
# quaternions, layout [..., (w, x, y, z)]            (rotations/quat.py:17)
# ----------------------------------------------------------------------------
def _wxyz(q):
    return q[..., 0], q[..., 1], q[..., 2], q[..., 3]


def _cols(q):
    """Keep-dim component views ([..., 1] each), as quat.py:350-351 slices them."""
    return q[..., 0:1], q[..., 1:2], q[..., 2:3], q[..., 3:4]


def _cross3(a, b):
    """quat.py:653-674 (_fast_cross): component-wise cross product."""
    ax, ay, az = a[..., 0:1], a[..., 1:2], a[..., 2:3]
    bx, by, bz = b[..., 0:1], b[..., 1:2], b[..., 2:3]
    return np.concatenate([ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx], axis=-1)


def quat_length(q):
    """quat.py:364-376: Euclidean norm over the last axis (dtype preserved)."""
    return np.linalg.norm(q, axis=-1)


def quat_normalize(q, eps: float = 1e-8):
    """quat.py:411-423: q / (|q| + eps) -- eps is added to the NORM."""
    return q / (quat_length(q)[..., np.newaxis] + eps)


def quat_conjugate(q):
    """quat.py:396-408: (w, -x, -y, -z)."""
    return np.concatenate((q[..., 0:1], -q[..., 1:]), axis=-1)


def quat_inverse(q):
    """quat.py:379-393: inverse == conjugate (unit quaternion assumed)."""
    return quat_conjugate(q)


def quat_mul(a, b):
    """quat.py:337-361: Hamilton product, term order kept exactly."""
    aw, ax, ay, az = _cols(a)
    bw, bx, by, bz = _cols(b)
    return np.concatenate(
        (
            aw * bw - ax * bx - ay * by - az * bz,
            aw * bx + bw * ax + ay * bz - az * by,
            aw * by + bw * ay + az * bx - ax * bz,
            aw * bz + bw * az + ax * by - ay * bx,
        ),
        axis=-1,
    )


def quat_mul_vec(q, v):
    """quat.py:320-334: t = 2 (q_xyz x v);  v' = v + w t + q_xyz x t."""
    u = q[..., 1:]
    t = 2.0 * _cross3(u, v)
    return v + q[..., 0][..., np.newaxis] * t + _cross3(u, t)


def quat_to_matrix(q):
    """quat.py:276-317: 12 products in the INPUT dtype, stored into a float64
    array (np.empty default dtype, quat.py:306).  No normalisation inside."""
    w, x, y, z = _wxyz(q)
    x2, y2, z2 = x + x, y + y, z + z
    xx, yy, zz = x * x2, y * y2, z * z2
    xy, xz, yz = x * y2, x * z2, y * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    m = np.empty(q.shape[:-1] + (3, 3))
    m[..., 0, 0] = 1.0 - (yy + zz)
    m[..., 0, 1] = xy - wz
    m[..., 0, 2] = xz + wy
    m[..., 1, 0] = xy + wz
    m[..., 1, 1] = 1.0 - (xx + zz)
    m[..., 1, 2] = yz - wx
    m[..., 2, 0] = xz - wy
    m[..., 2, 1] = yz + wx
    m[..., 2, 2] = 1.0 - (xx + yy)
    return m


def quat_from_matrix(m):
    """quat.py:85-156: four-branch (Shepperd-style) extraction selected by
    m22 < 0, m00 > m11, m00 < -m11, then quat.normalize."""
    m00, m01, m02 = m[..., 0, 0], m[..., 0, 1], m[..., 0, 2]
    m10, m11, m12 = m[..., 1, 0], m[..., 1, 1], m[..., 1, 2]
    m20, m21, m22 = m[..., 2, 0], m[..., 2, 1], m[..., 2, 2]

    def pack(a, b, c, d):
        return np.stack([a, b, c, d], axis=-1)

    cand_x = pack(m21 - m12, 1.0 + m00 - m11 - m22, m10 + m01, m02 + m20)   # quat.py:115-123
    cand_y = pack(m02 - m20, m10 + m01, 1.0 - m00 + m11 - m22, m21 + m12)   # quat.py:124-132
    cand_z = pack(m10 - m01, m02 + m20, m21 + m12, 1.0 - m00 - m11 + m22)   # quat.py:136-144
    cand_w = pack(1.0 + m00 + m11 + m22, m21 - m12, m02 - m20, m10 - m01)   # quat.py:145-153
    neg = np.where((m00 > m11)[..., np.newaxis], cand_x, cand_y)
    pos = np.where((m00 < -m11)[..., np.newaxis], cand_z, cand_w)
    return quat_normalize(np.where((m22 < 0.0)[..., np.newaxis], neg, pos))


# ----------------------------------------------------------------------------
# dual quaternions, layout [..., (w_r x_r y_r z_r | w_d x_d y_d z_d)]
#                                                    (rotations/dual_quat.py:4-9)
# ----------------------------------------------------------------------------
def dq_from_rotation_translation(rotations, translations):
    """dual_quat.py:12-36: q_d = 0.5 * ((0, t) (x) q_r); the pure quaternion is
    built in a float64 np.zeros buffer (dual_quat.py:32) so the result is
    float64 whatever the input dtype."""
    pure = np.zeros(translations.shape[:-1] + (4,))
    pure[..., 1:] = translations
    dual = 0.5 * quat_mul(pure, rotations)
    return np.concatenate((rotations, dual), axis=-1)


def dq_from_translation(translations):
    """dual_quat.py:39-59: identity rotation, dual part = t / 2 (float64)."""
    out = np.zeros(translations.shape[:-1] + (8,))
    out[..., 0:1] = 1
    out[..., 5:] = translations * 0.5
    return out


def dq_to_rotation_translation(dq):
    """dual_quat.py:62-83: R = real part; t = vector part of 2 q_d (x) conj(q_r).
    Returns (rotations, translations)."""
    dq = dq.copy()
    real, dual = dq[..., :4], dq[..., 4:]
    trans = (2 * quat_mul(dual, quat_conjugate(real)))[..., 1:]
    return real, trans


# ----------------------------------------------------------------------------
# skeleton ops                                          (ops/skeleton.py)
# ----------------------------------------------------------------------------
def fk(rot, global_pos, offsets, parents):
    """ops/skeleton.py:16-61.

    q^ = normalize(rot); local 4x4 = [R(q^) | offsets_i] with the root's
    translation replaced by global_pos; then for i = 1..J-1 IN INDEX ORDER
    G_i = G_parents[i] @ G_i as float64 4x4 products.  Returns
    (positions[..., J, 3], rotmats[..., J, 3, 3]) (float64).  parents[0] is
    never read.
    """
    n_joints = rot.shape[-2]
    xf = np.zeros(rot.shape[:-1] + (4, 4))                       # skeleton.py:44
    xf[..., :3, :3] = quat_to_matrix(quat_normalize(rot))        # :45
    xf[..., :3, 3] = offsets                                     # :46
    xf[..., 3, 3] = 1                                            # :47
    xf[..., 0, :3, 3] = global_pos                               # :49
    for joint in range(1, n_joints):                             # :51-58
        up = parents[joint]
        xf[..., joint, :, :] = np.matmul(xf[..., up, :, :], xf[..., joint, :, :])
    return xf[..., :3, 3], xf[..., :3, :3]


def to_root_dual_quat(rotations, global_pos, parents, offsets):
    """ops/skeleton.py:207-244.

    Root-centred accumulation in the INPUT dtype: joints whose parent is joint
    0 are left as they are (:236-237); the others compose with their parent's
    already-accumulated (R, t).  Joint 0 carries (r_0, global_pos).  No
    normalisation.  Joint count is taken from axis 1 (:228) like the reference.
    """
    assert (offsets[0] == np.zeros(3)).all()                     # :227
    n_joints = rotations.shape[1]                                # :228
    rots = rotations.copy()
    trans = np.tile(offsets, rots.shape[:-2] + (1, 1))           # :231
    trans[..., 0, :] = global_pos                                # :232
    for joint in range(1, n_joints):
        up = parents[joint]
        if up == 0:
            continue
        trans[..., joint, :] = quat_mul_vec(rots[..., up, :], trans[..., joint, :]) + trans[..., up, :]
        rots[..., joint, :] = quat_mul(rots[..., up, :], rots[..., joint, :])
    return dq_from_rotation_translation(rots, trans)             # :243


def from_root_dual_quat(dq, parents):
    """ops/skeleton.py:173-204.  Returns (translations, rotations) -- that
    order (:204), the docstring of the reference notwithstanding.

    Walks joints from last to first so a joint's parent is still expressed in
    root space when it is used (:194-203).
    """
    n_joints = dq.shape[1]                                       # :188
    rots, trans = dq_to_rotation_translation(dq.copy())          # :191
    for joint in range(n_joints - 1, 0, -1):
        up = parents[joint]
        if up == 0:
            continue
        back = quat_inverse(rots[..., up, :])
        trans[..., joint, :] = quat_mul_vec(back, trans[..., joint, :] - trans[..., up, :])
        rots[..., joint, :] = quat_mul(back, rots[..., joint, :])
    return trans, rots


def from_global_rotations(global_quats, parents):
    """ops/skeleton.py:64-93: local_i = conj(global_parents[i]) (x) global_i for
    i >= 1, root unchanged.  (SURVEY section 8f rank 1 -- 'next' row.)"""
    parents = np.asarray(parents)
    local = np.empty_like(global_quats)
    local[..., 0, :] = global_quats[..., 0, :]
    up = quat_inverse(global_quats[..., parents[1:], :])
    local[..., 1:, :] = quat_mul(up, global_quats[..., 1:, :])
    return local
