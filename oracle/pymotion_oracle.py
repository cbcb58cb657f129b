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
    # the rest of the quat / dual_quat surface
    "vec_normalize", "quat_from_angle_axis", "quat_from_scaled_angle_axis", "quat_from_euler", "quat_to_euler",
    "quat_to_angle_axis", "quat_to_scaled_angle_axis", "quat_unroll", "quat_slerp", "quat_from_to",
    "quat_from_to_axis", "dq_is_unit", "dq_normalize", "dq_unroll",
    # the in-repo consumers of fk
    "from_root_positions", "mirror",
    # SURVEY 8f rank 3 (ortho6d) and rank 4 (the consumers / helpers either side of fk)
    "ortho6d_from_quat", "ortho6d_from_matrix", "ortho6d_to_matrix", "ortho6d_to_quat", "center_of_mass",
    "human_center_of_mass", "interpolate_positions",
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


# ----------------------------------------------------------------------------
# the rest of the quaternion surface      (rotations/quat.py)
# dtype promotions of the reference are reproduced on purpose (noted per function)
# ----------------------------------------------------------------------------
_AXIS_INDEX = {"x": 0, "y": 1, "z": 2}


def vec_normalize(v, eps: float = 1e-8):
    """ops/vector.py:4-19: v / (|v| + eps)."""
    return v / (np.linalg.norm(v, axis=-1, keepdims=True) + eps)


def quat_from_angle_axis(angle, axis):
    """quat.py:24-40: (cos(a/2), sin(a/2) * axis); angle is [..., 1], axis [..., 3] unit."""
    half = angle / 2.0
    return np.concatenate((np.cos(half), np.sin(half) * axis), axis=-1)


def quat_from_scaled_angle_axis(scaledaxis):
    """quat.py:6-21: angle = |v|, axis = v / angle (0/0 = nan for the null vector, as in the reference)."""
    angle = np.linalg.norm(scaledaxis, axis=-1)[..., np.newaxis]
    return quat_from_angle_axis(angle, scaledaxis / angle)


def _order_indices(order):
    """'x'|'y'|'z' -> 0|1|2 per element (the reference does this with np.apply_along_axis, quat.py:72-81, :191-193)."""
    order = np.asarray(order)
    idx = np.full(order.shape, -1, dtype=np.int64)
    for ch, k in _AXIS_INDEX.items():
        idx[order == ch] = k
    if (idx < 0).any():
        raise KeyError("order entries must be 'x', 'y' or 'z'")
    return idx


def quat_from_euler(euler, order):
    """quat.py:43-82: q = q(e0 about order0) (x) q(e1 about order1) (x) q(e2 about order2).  The unit axes are
    INTEGER arrays in the reference (quat.py:65-69), so sin * axis -- and everything after it -- is float64
    even for float32 angles (cos / sin themselves are taken in the input dtype)."""
    assert euler.shape[:-1] == np.asarray(order).shape[:-1], "euler and order must have the same shape except for the last dimension"
    idx = _order_indices(order)
    eye = np.eye(3, dtype=np.int64)
    qs = [quat_from_angle_axis(euler[..., k:k + 1], eye[idx[..., k]]) for k in range(3)]
    return quat_mul(qs[0], quat_mul(qs[1], qs[2]))


def quat_to_euler(quaternions, order):
    """quat.py:159-227: intrinsic Euler angles through the half-angle sum / difference construction, every
    angle reduced with np.mod(., 2 pi).  `sign` is an int64 array, so the two terms multiplied by it are
    float64; the output buffer is float64 (np.empty default, quat.py:200)."""
    assert quaternions.shape[:-1] == np.asarray(order).shape[:-1], "quaternions and order must have the same shape except for the last dimension"
    idx = _order_indices(order)
    i, j, k = idx[..., 2:3], idx[..., 1:2], idx[..., 0:1]
    sign = (i - j) * (j - k) * (k - i) // 2  # +1 even permutation, -1 odd
    w = quaternions[..., 0:1]
    qi = np.take_along_axis(quaternions, i + 1, axis=-1)
    qj = np.take_along_axis(quaternions, j + 1, axis=-1)
    qk = np.take_along_axis(quaternions, k + 1, axis=-1)
    a = w - qj
    b = qi + qk * sign
    c = qj + w
    d = qk * sign - qi
    euler = np.empty(quaternions.shape[:-1] + (3,))
    euler[..., 1:2] = (2 * np.arctan2(np.hypot(c, d), np.hypot(a, b))) - (np.pi / 2)
    half_sum, half_diff = np.arctan2(b, a), np.arctan2(d, c)
    euler[..., 2:3] = half_sum - half_diff
    euler[..., 0:1] = (half_sum + half_diff) * sign
    return np.mod(euler, 2 * np.pi)


def quat_to_angle_axis(quaternions):
    """quat.py:247-273: angle = 2 acos(clip(w)); axis = xyz / sqrt(clip(1 - w^2)) where that root exceeds 1e-8,
    zero elsewhere.  Returns (angle [..., 1], axis [..., 3])."""
    w, xyz = quaternions[..., 0], quaternions[..., 1:]
    angle = 2 * np.arccos(np.clip(w, -1.0, 1.0))
    s = np.sqrt(np.clip(1.0 - w * w, 0.0, 1.0))
    axis = np.zeros_like(xyz)
    ok = s > 1e-8
    if ok.any():
        axis[ok] = xyz[ok] / np.expand_dims(s[ok], axis=-1)
    return angle[..., np.newaxis], axis


def quat_to_scaled_angle_axis(quaternions):
    """quat.py:230-244."""
    angle, axis = quat_to_angle_axis(quaternions)
    return angle * axis


def _unroll_signs(real, axis):
    """Sign (+1 / -1) each entry along `axis` ends up with in quat.py:450-462 / dual_quat.py:155-167: entry i is
    negated iff its dot product with the ALREADY UNROLLED entry i-1 is < 0 (strictly)."""
    r = np.moveaxis(real, axis, 0)
    sign = np.ones(r.shape[:-1], dtype=real.dtype)
    for t in range(1, r.shape[0]):
        d0 = np.sum(r[t] * (sign[t - 1][..., np.newaxis] * r[t - 1]), axis=-1)
        sign[t] = np.where(d0 < -d0, -1.0, 1.0)
    return np.moveaxis(sign, 0, axis)


def quat_unroll(quaternions, axis):
    """quat.py:426-462.  The reference flips IN PLACE through a swapaxes view (its input is modified); the
    restatement returns a new array with the same values."""
    return quaternions * _unroll_signs(quaternions, axis)[..., np.newaxis]


def quat_slerp(q0, q1, t, shortest: bool = True):
    """quat.py:465-501: note the 1e-6 added to every COMPONENT of q2 before its norm is taken (:499)."""
    dot = np.sum(q0 * q1, axis=-1, keepdims=True)
    flip = np.logical_and(shortest, dot < 0)
    q1 = np.where(flip, -q1, q1)
    dot = np.clip(np.where(flip, -dot, dot), -1, 1)
    theta = np.arccos(dot) * t
    q2 = q1 - q0 * dot
    q2 = q2 / np.linalg.norm(q2 + 0.000001, axis=-1, keepdims=True)
    return np.cos(theta) * q0 + np.sin(theta) * q2


def _from_to_setup(v1, v2, normalize_input):
    assert v1.shape[-1] == 3 and v2.shape[-1] == 3, "Input vectors must have shape [..., 3]"
    assert v1.shape == v2.shape, "Input vectors must have the same shape"
    a, b = (vec_normalize(v1), vec_normalize(v2)) if normalize_input else (v1, v2)
    if v1.ndim == 1:
        a, b = a[np.newaxis, :], b[np.newaxis, :]
    return a, b, np.cross(a, b), np.sum(a * b, axis=-1, keepdims=True)


def quat_from_to(v1, v2, normalize_input: bool = True):
    """quat.py:504-576: half-angle quaternion about normalize(v1 x v2); identity where np.isclose(dot, 1);
    where np.isclose(dot, -1) a half turn about normalize(v1 x e), e = y if |v1.x| is close to 1 else x."""
    a, b, cross, dot = _from_to_setup(v1, v2, normalize_input)
    rot = np.concatenate([np.sqrt((1 + dot) * 0.5), quat_normalize(cross) * np.sqrt((1 - dot) * 0.5)], axis=-1)
    rot[np.isclose(dot, 1.0)[..., 0]] = [1.0, 0.0, 0.0, 0.0]
    anti = np.isclose(dot, -1.0)[..., 0]
    if np.any(anti):
        va = a[anti]
        ortho = np.empty_like(va)
        along_x = np.isclose(np.abs(va[..., 0]), 1.0)
        ortho[along_x] = np.array([0.0, 1.0, 0.0])
        ortho[~along_x] = np.array([1.0, 0.0, 0.0])
        ax = quat_normalize(np.cross(va, ortho))
        rot[anti] = np.concatenate([np.zeros_like(ax[..., :1]), ax], axis=-1)
    return rot[0] if v1.ndim == 1 else rot


def quat_from_to_axis(v1, v2, rot_axis, normalize_input: bool = True):
    """quat.py:579-650: same half angle about the GIVEN axis, signed by sign((v1 x v2) . axis); identity where
    np.isclose(dot, 1); (0, axis) where np.isclose(dot, -1)."""
    assert v1.shape == rot_axis.shape, "Input vectors and rotation axis must have the same shape"
    if rot_axis.ndim == 1:
        rot_axis = rot_axis[np.newaxis, :]
    a, b, cross, dot = _from_to_setup(v1, v2, normalize_input)
    s = np.sqrt((1 - dot) * 0.5) * np.sign(np.sum(cross * rot_axis, axis=-1, keepdims=True))
    rot = np.concatenate([np.sqrt((1 + dot) * 0.5), rot_axis * s], axis=-1)
    rot[np.isclose(dot, 1.0)[..., 0]] = [1.0, 0.0, 0.0, 0.0]
    anti = np.isclose(dot, -1.0)[..., 0]
    if np.any(anti):
        rot[anti] = np.concatenate([np.zeros_like(rot_axis[anti][..., :1]), rot_axis[anti]], axis=-1)
    return rot[0] if v1.ndim == 1 else rot


# ----------------------------------------------------------------------------
# the rest of the dual-quaternion surface               (rotations/dual_quat.py)
# ----------------------------------------------------------------------------
def dq_is_unit(dq, atol: float = 1e-03) -> bool:
    """dual_quat.py:118-136: ONE bool for the whole array -- every real part has unit norm (np.isclose
    defaults) and is orthogonal to its dual part (|dot| <= atol); all-zero real parts count as unit."""
    real, dual = dq[..., :4], dq[..., 4:]
    n2 = np.sum(real * real, axis=-1)
    if np.isclose(n2, 0).all():
        return True
    return bool(np.isclose(n2, 1).all() and np.isclose(np.sum(real * dual, axis=-1), 0, atol=atol).all())


def dq_normalize(dq):
    """dual_quat.py:86-115: both parts divided by |real|; if the WHOLE array is then not unit (dq_is_unit) the
    component of the dual part along the real part is removed from every element."""
    real, dual = dq[..., :4], dq[..., 4:]
    norm = np.linalg.norm(real, axis=-1)
    rn, dn = real / norm[..., np.newaxis], dual / norm[..., np.newaxis]
    if not dq_is_unit(np.concatenate((rn, dn), axis=-1)):
        dn = dn - rn * (np.sum(real * dual, axis=-1) / (norm * norm))[..., np.newaxis]
    return np.concatenate((rn, dn), axis=-1)


def dq_unroll(dq, axis):
    """dual_quat.py:139-167: quat unroll decided on the REAL part, the flip applied to all eight numbers."""
    return dq * _unroll_signs(dq[..., :4], axis)[..., np.newaxis]


# ----------------------------------------------------------------------------
# fk consumers                         (ops/skeleton.py)
# ----------------------------------------------------------------------------
def from_root_positions(positions, parents, offsets):
    """ops/skeleton.py:96-170.  Joints are visited in index order; before EVERY alignment the current
    rotations go through a full fk (root at the origin) and quat.from_matrix, exactly like the reference:
      first child c of j   : rot_j = from_to(G_j^-1 (p_c - p_j), G_j^-1 (P_c - P_j))
      further children g   : rot_j = rot_j (x) from_to_axis(G_j^-1 (p_g - p_j), G_j^-1 (P_g - P_j),
                                                           G_j^-1 normalize(P_c - P_j))
    with p the fk pose under the rotations chosen so far, P the target positions and G_j the global rotation
    of j in that pose.  Result float64 (the rotations start as a float64 identity, :129)."""
    parents = np.asarray(parents)
    n_frames, n_joints = positions.shape[0], parents.shape[0]
    kids = [[] for _ in range(n_joints)]
    for i in range(1, n_joints):
        kids[parents[i]].append(i)
    rotations = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (n_frames, n_joints, 1))
    origin = np.zeros((1, 3))

    def pose():
        pos, rotm = fk(rotations, origin, offsets, parents)
        return pos, quat_from_matrix(rotm)

    for j, ks in enumerate(kids):
        if not ks:
            continue
        pos, grot = pose()
        back = quat_inverse(grot[:, j])
        c = ks[0]
        rest = quat_mul_vec(back, pos[:, c] - pos[:, j])
        pred = quat_mul_vec(back, positions[:, c] - positions[:, j])
        rotations[:, j] = quat_from_to(rest, pred)
        for g in ks[1:]:
            pos, grot = pose()
            back = quat_inverse(grot[:, j])
            rest_g = quat_mul_vec(back, pos[:, g] - pos[:, j])
            pred_g = quat_mul_vec(back, positions[:, g] - positions[:, j])
            roll_axis = quat_mul_vec(back, vec_normalize(positions[:, c] - positions[:, j]))
            rotations[:, j] = quat_mul(rotations[:, j], quat_from_to_axis(rest_g, pred_g, roll_axis))
    return rotations


_MIRROR = {"X": (0, (2, 3)), "Y": (1, (1, 3)), "Z": (2, (1, 2))}  # axis -> (vector index, quaternion indices negated)


def _mirror_all(local_rotations, global_translation, parents, offsets, end_sites, axis):
    """ops/skeleton.py:347-418 (_true_mirror): offsets / end sites / root translation reflected, global
    quaternions (fk with the reflected offsets, root at the origin -> from_matrix) get two components negated
    and go back to local space."""
    if axis not in _MIRROR:
        raise ValueError("Invalid axis. Choose 'X', 'Y', or 'Z'")
    vi, (qa, qb) = _MIRROR[axis]
    offsets = offsets.copy()
    offsets[:, vi] = -offsets[:, vi]
    if end_sites is not None:
        end_sites = end_sites.copy()
        end_sites[:, vi] = -end_sites[:, vi]
    global_translation = global_translation.copy()
    global_translation[..., vi] = -global_translation[..., vi]
    _, rotm = fk(local_rotations, np.zeros_like(global_translation), offsets, parents)
    gq = quat_from_matrix(rotm)
    gq[..., qa] = -gq[..., qa]
    gq[..., qb] = -gq[..., qb]
    return from_global_rotations(gq, parents), global_translation, offsets, end_sites


def mirror(local_rotations, global_translation, parents, offsets, end_sites=None, joints_mapping=None, mode="all",
           axis="X"):
    """ops/skeleton.py:247-344.  Unlike the reference ('symmetry' mode negates the caller's global_translation
    in place, :321) the restatement never modifies its inputs; returned values are the same."""
    parents = np.asarray(parents)
    if mode == "all":
        return _mirror_all(local_rotations, global_translation, parents, offsets, end_sites, axis)
    if mode == "symmetry":
        if joints_mapping is None:
            raise ValueError("joints_mapping must be provided for mode 'symmetry'")
        if len(joints_mapping) != len(parents):
            raise ValueError("joints_mapping must have the same length as the number of joints")
        if axis not in _MIRROR:
            raise ValueError("Invalid axis. Choose 'X', 'Y', or 'Z'")
        vi, (qa, qb) = _MIRROR[axis]
        _, rotm = fk(local_rotations, np.zeros_like(global_translation), offsets, parents)
        gq = quat_from_matrix(rotm)[..., np.asarray(joints_mapping), :]
        gq[..., qa] = -gq[..., qa]
        gq[..., qb] = -gq[..., qb]
        moved = global_translation.copy()
        moved[..., vi] = -moved[..., vi]
        return from_global_rotations(gq, parents), moved, offsets, end_sites
    if mode == "positions":
        rots, moved, mirrored_offsets, _ = _mirror_all(local_rotations, global_translation, parents, offsets, end_sites, axis)
        pos, _ = fk(rots, moved, mirrored_offsets, parents)
        pos = pos - pos[..., 0:1, :]
        return from_root_positions(pos, parents, offsets), moved, offsets, end_sites
    raise ValueError("Invalid mode. Choose 'symmetry', 'all', or 'positions'")


# ----------------------------------------------------------------------------
# 6-D rotation representation                           (rotations/ortho6d.py)
# [..., 3, 2] = the first two COLUMNS of the rotation matrix
# ----------------------------------------------------------------------------
def ortho6d_from_matrix(rotmats):
    """ortho6d.py:31-47 (a view of the input, like the reference)."""
    return rotmats[..., :2]


def ortho6d_from_quat(quaternions):
    """ortho6d.py:14-28."""
    return ortho6d_from_matrix(quat_to_matrix(quaternions))


def ortho6d_to_matrix(ortho6d):
    """ortho6d.py:67-90: Gram-Schmidt on the two columns (plain norms, no eps), third column = cross product."""
    a, b = ortho6d[..., 0], ortho6d[..., 1]
    c1 = a / np.linalg.norm(a, axis=-1, keepdims=True)
    c2 = b - np.sum(c1 * b, axis=-1)[..., np.newaxis] * c1
    c2 = c2 / np.linalg.norm(c2, axis=-1, keepdims=True)
    c3 = np.cross(c1, c2, axis=-1)
    rows = np.concatenate([c1, c2, c3], axis=-1).reshape(*ortho6d.shape[:-2], 3, 3)
    return np.swapaxes(rows, -2, -1)


def ortho6d_to_quat(ortho6d):
    """ortho6d.py:50-64."""
    return quat_from_matrix(ortho6d_to_matrix(ortho6d))


# ----------------------------------------------------------------------------
# consumers of fk positions                 (ops/center_of_mass.py, ops/time.py)
# ----------------------------------------------------------------------------
def center_of_mass(joints, weights):
    """center_of_mass.py:52-68: sum over the joint axis of joints * weights."""
    return np.sum(joints * weights[..., np.newaxis], axis=-2)


def human_center_of_mass(joints_spine, joints_left_arm, joints_right_arm, joints_left_leg, joints_right_leg):
    """center_of_mass.py:4-49: spine 60 %, each arm 5 %, each leg 15 %, spread evenly over the joints of the part."""
    parts = (joints_spine, joints_left_arm, joints_right_arm, joints_left_leg, joints_right_leg)
    shares = (0.6, 0.05, 0.05, 0.15, 0.15)
    weights = np.array([w for part, share in zip(parts, shares) for w in [share / part.shape[-2]] * part.shape[-2]])
    return center_of_mass(np.concatenate(parts, axis=-2), weights)


def interpolate_positions(sample_times, original_times, positions, axis, method="linear"):
    """time.py:4-66: linear interpolation along `axis`; the bracketing interval comes from np.searchsorted
    (left) clamped to [0, T-2], so samples outside the original range extrapolate from the end intervals."""
    assert method == "linear", "Only linear interpolation is supported yet."
    assert positions.shape[axis] == original_times.shape[0], (
        "Wrong shape of data. Positions along the axis dimension must be equal to the length of original_times.")
    idx = np.minimum(np.maximum(np.searchsorted(original_times, sample_times) - 1, 0), original_times.shape[0] - 2)
    w = (sample_times - original_times[idx]) / (original_times[idx + 1] - original_times[idx])
    lo = np.take(positions, idx, axis=axis)
    hi = np.take(positions, idx + 1, axis=axis)
    shape = [1] * positions.ndim
    shape[axis] = len(sample_times)
    # the reference broadcasts the weights as weights[..., np.newaxis] against positions indexed on `axis`
    # (time.py:61-64), which lines them up with `axis` only when it is the second-to-last axis or the array is
    # [T, 3]; the restatement aligns them with `axis` explicitly (same result wherever the reference works)
    wb = w.reshape(shape)
    out = (1 - wb) * lo
    out += wb * hi
    return out


# ----------------------------------------------------------------------------
# io/bvh.py: the numeric part of BVH.get_data
# ----------------------------------------------------------------------------
def bvh_get_data_rotations(rotations_deg, rot_order):
    """io/bvh.py:352-359: rots = normalize(unroll(from_euler(np.radians(rotations), order tiled over the frames),
    axis=0)).  rotations_deg [n_frames, n_joints, 3] in degrees, rot_order [n_joints, 3] of 'x' | 'y' | 'z'."""
    order = np.tile(rot_order, (rotations_deg.shape[0], 1, 1))
    rots = quat_unroll(quat_from_euler(np.radians(rotations_deg), order), axis=0)
    return quat_normalize(rots)
