#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REAL reference (run in the build
container only: needs /root/reference, which does not exist on the GPU box).

    python oracle/gen_golden.py

Every array written here is an output of the unmodified reference
(`/root/reference/pymotion`, NumPy path) on seeded inputs that are stored next
to it, plus the hand-written golden arrays transcribed from the reference's own
tests (cited below).  The fixtures pin `oracle/pymotion_oracle.py`
(tests/test_oracle_golden.py) and are what the `-m gpu` parity tests compare
the CUDA path against.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

import pymotion.ops.center_of_mass as ref_com  # noqa: E402
import pymotion.ops.skeleton as ref_sk  # noqa: E402
import pymotion.ops.time as ref_time  # noqa: E402
import pymotion.ops.vector as ref_vec  # noqa: E402
import pymotion.rotations.ortho6d as ref_o6  # noqa: E402
import pymotion.rotations.dual_quat as ref_dq  # noqa: E402
import pymotion.rotations.quat as ref_q  # noqa: E402

from pymotion_b200.topologies import TOPOLOGIES, synth_numpy  # noqa: E402

OUT = os.path.join(REPO, "tests", "golden")


def rx(a):
    return [[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]]


def ry(a):
    return [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]


def rz(a):
    return [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]


def chain3():
    """The 3-joint chain every skeleton test of the reference uses
    (ops/tests/test_skeleton.py:26-34, :237-245)."""
    offsets = np.array([[0, 0, 0], [0, 0, 1], [0, 0, 2]], dtype=np.float32)
    parents = np.array([0, 0, 1])
    gpos = np.array([[0, 0, 0], [1, 1, 1]], dtype=np.float32)
    ident = np.tile(np.array([1, 0, 0, 0], dtype=np.float32), (2, 3, 1))
    rotm = np.array(
        [[rx(np.pi / 2), ry(np.pi / 2), rz(np.pi / 2)], [ry(np.pi / 4), rz(np.pi / 4), rx(np.pi / 4)]]
    )  # test_skeleton.py:312-325 (float64)
    return offsets, parents, gpos, ident, rotm


def gen_fk():
    d = {}
    offsets, parents, gpos, ident, rotm = chain3()
    d["chain3/offsets"], d["chain3/parents"], d["chain3/gpos"] = offsets, parents, gpos
    # identity case + its hand-written golden (test_skeleton.py:247-272)
    d["chain3_ident/rot"] = ident
    p, r = ref_sk.fk(ident, gpos, offsets, parents)
    d["chain3_ident/pos"], d["chain3_ident/rotm"] = np.ascontiguousarray(p), np.ascontiguousarray(r)
    d["chain3_ident/hand_pos"] = np.array(
        [[[0, 0, 0], [0, 0, 1], [0, 0, 3]], [[1, 1, 1], [1, 1, 2], [1, 1, 4]]], dtype=np.float32
    )  # :253-258
    # per-frame offsets accepted (:267)
    p, r = ref_sk.fk(ident, gpos, np.tile(offsets, (2, 1, 1)), parents)
    d["chain3_ident/pos_pf"], d["chain3_ident/rotm_pf"] = np.ascontiguousarray(p), np.ascontiguousarray(r)
    # rotated case (:312-384): quaternions come from quat.from_matrix of float64 matrices
    q = ref_q.from_matrix(rotm)
    d["chain3_rot/rot"] = q
    p, r = ref_sk.fk(q, gpos, offsets, parents)
    d["chain3_rot/pos"], d["chain3_rot/rotm"] = np.ascontiguousarray(p), np.ascontiguousarray(r)
    d["chain3_rot/hand_pos"] = np.array(
        [[[0, 0, 0], [0, -1, 0], [2, -1, 0]], [[1, 1, 1], [1.707107, 1, 1.707107], [3.12132, 1, 3.12132]]],
        dtype=np.float32,
    )  # :326-331
    d["chain3_rot/hand_rotm"] = np.array(
        [
            [[[1, 0, 0], [0, 0, -1], [0, 1, 0]], [[0, 0, 1], [1, 0, 0], [0, 1, 0]], [[0, 0, 1], [0, -1, 0], [1, 0, 0]]],
            [
                [[0.7071068, 0, 0.7071068], [0, 1, 0], [-0.7071068, 0, 0.7071068]],
                [[0.5, -0.5, 0.7071068], [0.7071068, 0.7071068, 0], [-0.5, 0.5, 0.7071068]],
                [[0.5, 0.1464466, 0.8535535], [0.7071069, 0.5, -0.5], [-0.5, 0.8535535, 0.1464465]],
            ],
        ],
        dtype=np.float32,
    )  # :332-369
    # N-D leading dims (:386-404)
    qn = np.tile(q, (4, 3, 2, 1, 1))
    gn = np.tile(gpos, (4, 3, 2, 1))
    p, r = ref_sk.fk(qn, gn, offsets, parents)
    d["chain3_nd/rot"], d["chain3_nd/gpos"] = qn, gn
    d["chain3_nd/pos"], d["chain3_nd/rotm"] = np.ascontiguousarray(p), np.ascontiguousarray(r)

    # seeded differential cases on realistic skeletons (gap named in SURVEY 8c)
    for name, frames, seed in (("body22", 37, 11), ("smplh52", 19, 12), ("deep65", 13, 13)):
        par = np.asarray(TOPOLOGIES[name])
        rot, gp, off = synth_numpy(frames, par, seed=seed)
        rot = rot * np.random.default_rng(seed).uniform(0.25, 4.0, size=rot.shape[:-1] + (1,)).astype(np.float32)
        rot[0, 1] = 0.0  # zero quaternion -> identity after q/(0+eps)   (SURVEY 8a notes)
        p, r = ref_sk.fk(rot, gp, off, par)
        d[f"{name}/parents"], d[f"{name}/rot"], d[f"{name}/gpos"], d[f"{name}/offsets"] = par, rot, gp, off
        d[f"{name}/pos"], d[f"{name}/rotm"] = np.ascontiguousarray(p), np.ascontiguousarray(r)
        # per-frame offsets, and offsets[0] != 0 (ignored by fk, skeleton.py:49)
        off_pf = np.tile(off, (frames, 1, 1)) * np.linspace(0.5, 1.5, frames, dtype=np.float32)[:, None, None]
        off_pf[:, 0] = 7.0
        p, r = ref_sk.fk(rot, gp, off_pf, par)
        d[f"{name}/offsets_pf"] = off_pf
        d[f"{name}/pos_pf"], d[f"{name}/rotm_pf"] = np.ascontiguousarray(p), np.ascontiguousarray(r)
        # float64 inputs
        p, r = ref_sk.fk(rot.astype(np.float64), gp.astype(np.float64), off.astype(np.float64), par)
        d[f"{name}/pos_f64"], d[f"{name}/rotm_f64"] = np.ascontiguousarray(p), np.ascontiguousarray(r)
    # unbatched [J,4] input, parents[0] = -1, single joint
    par = np.asarray(TOPOLOGIES["body22"])
    rot, gp, off = synth_numpy(1, par, seed=21)
    par_m1 = par.copy()
    par_m1[0] = -1
    p, r = ref_sk.fk(rot[0], gp[0], off, par_m1)
    d["unbatched/parents"], d["unbatched/rot"], d["unbatched/gpos"], d["unbatched/offsets"] = par_m1, rot[0], gp[0], off
    d["unbatched/pos"], d["unbatched/rotm"] = np.ascontiguousarray(p), np.ascontiguousarray(r)
    rot1, gp1, off1 = synth_numpy(5, [0], seed=22)
    p, r = ref_sk.fk(rot1, gp1, off1, np.array([0]))
    d["single/rot"], d["single/gpos"], d["single/offsets"] = rot1, gp1, off1
    d["single/pos"], d["single/rotm"] = np.ascontiguousarray(p), np.ascontiguousarray(r)
    # broadcast global_pos [1,3] against [F,J,4] (how from_root_positions calls fk, skeleton.py:134-139)
    rot, gp, off = synth_numpy(9, par, seed=23)
    p, r = ref_sk.fk(rot, np.zeros((1, 3)), off, par)
    d["bcast/rot"], d["bcast/offsets"] = rot, off
    d["bcast/pos"], d["bcast/rotm"] = np.ascontiguousarray(p), np.ascontiguousarray(r)
    np.savez_compressed(os.path.join(OUT, "fk.npz"), **d)
    return d


def gen_dq():
    d = {}
    offsets, parents, gpos, ident, rotm = chain3()
    d["chain3/offsets"], d["chain3/parents"], d["chain3/gpos"] = offsets, parents, gpos
    for tag, q in (("chain3_ident", ident), ("chain3_rot", ref_q.from_matrix(rotm))):
        dq = ref_sk.to_root_dual_quat(q, gpos, parents, offsets)
        t, r = ref_sk.from_root_dual_quat(dq, parents)
        rr, tt = ref_dq.to_rotation_translation(dq)
        d[f"{tag}/rot"], d[f"{tag}/dq"] = q, dq
        d[f"{tag}/back_trans"], d[f"{tag}/back_rot"] = np.ascontiguousarray(t), np.ascontiguousarray(r)
        d[f"{tag}/root_rot"], d[f"{tag}/root_trans"] = np.ascontiguousarray(rr), np.ascontiguousarray(tt)
    # hand-written goldens: root-space translations (test_skeleton.py:42-47, :123-128)
    d["chain3_ident/hand_root_trans"] = np.array(
        [[[0, 0, 0], [0, 0, 1], [0, 0, 3]], [[1, 1, 1], [0, 0, 1], [0, 0, 3]]], dtype=np.float32
    )
    d["chain3_rot/hand_root_trans"] = np.array(
        [[[0, 0, 0], [0, 0, 1], [2, 0, 1]], [[1, 1, 1], [0, 0, 1], [0, 0, 3]]], dtype=np.float32
    )  # joint 0 rows are compared against global_pos in the reference test (:171)
    d["chain3_rot/hand_root_rotm"] = np.array(
        [
            [[[1, 0, 0], [0, 0, -1], [0, 1, 0]], [[0, 0, 1], [0, 1, 0], [-1, 0, 0]], [[0, 0, 1], [1, 0, 0], [0, 1, 0]]],
            [
                [[0.7071068, 0, 0.7071068], [0, 1, 0], [-0.7071068, 0, 0.7071068]],
                [[0.7071068, -0.7071068, 0], [0.7071068, 0.7071068, 0], [0, 0, 1]],
                [[0.7071068, -0.5, 0.5], [0.7071068, 0.5, -0.5], [0, 0.7071068, 0.7071068]],
            ],
        ],
        dtype=np.float32,
    )  # :129-166
    for name, frames, seed in (("body22", 37, 31), ("smplh52", 19, 32), ("deep65", 13, 33)):
        par = np.asarray(TOPOLOGIES[name])
        rot, gp, off = synth_numpy(frames, par, seed=seed)
        # to_root_dual_quat does NOT normalise: feed slightly non-unit quaternions too
        rot = rot * np.random.default_rng(seed).uniform(0.9, 1.1, size=rot.shape[:-1] + (1,)).astype(np.float32)
        dq = ref_sk.to_root_dual_quat(rot, gp, par, off)
        t, r = ref_sk.from_root_dual_quat(dq, par)
        t32, r32 = ref_sk.from_root_dual_quat(dq.astype(np.float32), par)
        d[f"{name}/parents"], d[f"{name}/rot"], d[f"{name}/gpos"], d[f"{name}/offsets"] = par, rot, gp, off
        d[f"{name}/dq"] = dq
        d[f"{name}/back_trans"], d[f"{name}/back_rot"] = np.ascontiguousarray(t), np.ascontiguousarray(r)
        d[f"{name}/back_trans_f32in"], d[f"{name}/back_rot_f32in"] = np.ascontiguousarray(t32), np.ascontiguousarray(r32)
    # from_global_rotations (8f rank 1)
    par = np.asarray(TOPOLOGIES["body22"])
    rot, _, _ = synth_numpy(11, par, seed=41)
    d["fgr/parents"], d["fgr/global"] = par, rot
    d["fgr/local"] = ref_sk.from_global_rotations(rot, par)
    np.savez_compressed(os.path.join(OUT, "dq.npz"), **d)
    return d


def gen_quat():
    d = {}
    # hand-written samples shared by test_quat.py:205-211, :245-251, :279-285
    qa = np.array([[0.70710678, 0.70710678, 0, 0], [0.92387953, 0, 0.38268343, 0], [0, 0, 0, 1]])
    qb = np.array([[0, 0, 0, 1], [0.70710678, 0.70710678, 0, 0], [0.92387953, 0, 0.38268343, 0]])
    d["hand/qa"], d["hand/qb"] = qa, qb
    d["hand/mul_ab"] = np.array(
        [[0, 0, -0.70710678, 0.70710678], [0.65328148, 0.65328148, 0.27059805, -0.27059805], [0, -0.3826834, 0, 0.92387953]]
    )  # test_quat.py:296-302
    d["hand/mul_ba"] = np.array(
        [[0, 0, 0.70710678, 0.70710678], [0.65328148, 0.65328148, 0.27059805, 0.27059805], [0, 0.3826834, 0, 0.92387953]]
    )  # :303-309
    d["hand/v"] = np.array([[0, 2, 0], [0, 0, -4], [1, 2, 0]], dtype=np.float64)  # :253
    d["hand/mul_vec"] = np.array([[0, 0, 2], [-2.828427, 0, -2.828427], [-1, -2, 0]])  # :257-263
    d["hand/matrix"] = np.array(
        [
            [[1, 0, 0], [0, 0, -1], [0, 1, 0]],
            [[0.70710678, 0, 0.70710678], [0, 1, 0], [-0.70710678, 0, 0.70710678]],
            [[-1, 0, 0], [0, -1, 0], [0, 0, 1]],
        ]
    )  # :214-232
    rng = np.random.default_rng(51)
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        q0 = rng.standard_normal((3, 41, 4)).astype(dt)
        q1 = rng.standard_normal((3, 41, 4)).astype(dt)
        v = rng.standard_normal((3, 41, 3)).astype(dt)
        qu = (q0 / np.linalg.norm(q0, axis=-1, keepdims=True)).astype(dt)
        d[f"{tag}/q0"], d[f"{tag}/q1"], d[f"{tag}/v"], d[f"{tag}/qu"] = q0, q1, v, qu
        d[f"{tag}/mul"] = ref_q.mul(q0, q1)
        d[f"{tag}/mul_bcast"] = ref_q.mul(q0[:, :1], q1)  # broadcasting over the joint axis
        d[f"{tag}/mul_vec"] = ref_q.mul_vec(q0, v)
        d[f"{tag}/length"] = ref_q.length(q0)
        d[f"{tag}/normalize"] = ref_q.normalize(q0)
        d[f"{tag}/normalize_eps"] = ref_q.normalize(q0, eps=1e-2)
        d[f"{tag}/conjugate"] = ref_q.conjugate(q0)
        d[f"{tag}/inverse"] = ref_q.inverse(q0)
        d[f"{tag}/to_matrix"] = ref_q.to_matrix(q0)
        m = ref_q.to_matrix(qu)
        d[f"{tag}/unit_matrix"] = m
        d[f"{tag}/from_matrix"] = ref_q.from_matrix(m.astype(dt))
        t = rng.standard_normal((3, 41, 3)).astype(dt)
        dq = ref_dq.from_rotation_translation(qu, t)
        r2, t2 = ref_dq.to_rotation_translation(dq)
        d[f"{tag}/t"], d[f"{tag}/dq"] = t, dq
        d[f"{tag}/dq_rot"], d[f"{tag}/dq_trans"] = np.ascontiguousarray(r2), np.ascontiguousarray(t2)
        d[f"{tag}/dq_from_translation"] = ref_dq.from_translation(t)
    # all four from_matrix branches (quat.py:111-155): rotations by ~pi about x, y, z and identity-ish
    axes = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 1, 1]], dtype=np.float64)
    axes /= np.linalg.norm(axes, axis=-1, keepdims=True)
    ang = np.array([3.0, 3.1, 2.9, 0.3])[:, None]
    qbr = np.concatenate([np.cos(ang / 2), np.sin(ang / 2) * axes], axis=-1)
    mbr = ref_q.to_matrix(qbr)
    d["branches/matrix"], d["branches/quat"] = mbr, ref_q.from_matrix(mbr)
    np.savez_compressed(os.path.join(OUT, "quat.npz"), **d)
    return d


def gen_quat_ext():
    """the rest of quat.py / dual_quat.py.  Seeded inputs -> outputs of the real reference,
    plus the hand-written samples of rotations/tests/test_quat.py (cited)."""
    d = {}
    rng = np.random.default_rng(77)
    orders6 = np.array([list(p) for p in ("xyz", "xzy", "yxz", "yzx", "zxy", "zyx")])
    # hand-written samples
    d["hand/aa_axis"] = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float64)  # test_quat.py:55
    d["hand/aa_angle"] = np.array([0, np.pi / 2, np.pi / 4, np.pi])[..., np.newaxis]               # :57
    d["hand/aa_quat"] = np.array([[1, 0, 0, 0], [0.70710678, 0.70710678, 0, 0], [0.92387953, 0, 0.38268343, 0], [0, 0, 0, 1]])  # :61-68
    d["hand/saa"] = np.array([[np.pi / 2, 0, 0], [0, np.pi / 4, 0], [0, 0, np.pi]])                 # :101
    d["hand/saa_quat"] = np.array([[0.70710678, 0.70710678, 0, 0], [0.92387953, 0, 0.38268343, 0], [0, 0, 0, 1]])  # :105-111
    d["hand/euler"] = np.array([[np.pi / 2, 0, 0], [0, np.pi / 4, 0], [0, 0, np.pi]])               # :150
    d["hand/euler_order"] = np.array([["x", "y", "z"], ["x", "z", "y"], ["y", "x", "z"]])           # :152
    d["hand/euler_quat"] = np.array([[0.70710678, 0.70710678, 0, 0], [0.92387953, 0, 0, 0.38268343], [0, 0, 0, 1]])  # :155-161
    # slerp samples, test_quat.py:393-452
    q1 = ref_q.from_angle_axis(np.array([[0], [np.pi / 2], [np.pi]]), np.array([[0, 0, 1], [0, 0, 1], [0, 0, 1]]))
    q2 = ref_q.from_angle_axis(np.array([[np.pi / 2], [0], [np.pi]]), np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0]]))
    d["hand/slerp_q1"], d["hand/slerp_q2"] = q1, q2
    d["hand/slerp_t"] = np.array([[0], [0.5], [1]])
    d["hand/slerp_t2"] = np.array([[-1], [2], [0.25]])
    d["hand/slerp_gt"] = np.array([[1, 0, 0, 0], [0.92387953, 0, 0, 0.38268343], [0, 0, 1, 0]])
    d["hand/slerp_gt2"] = np.array([[0.70710678, 0, 0, -0.70710678], [0.70710678, 0, 0, -0.70710678], [0, 0, 0.38268343, 0.92387953]])
    d["hand/slerp_gt3"] = np.array([[0.83146961, 0, 0, 0.5555702], [0.98078528, 0, 0, 0.1950903], [0, 0, 0.9238795, 0.3826834]])  # t = 0.75
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        n = (3, 29)
        axis = rng.standard_normal(n + (3,))
        axis = (axis / np.linalg.norm(axis, axis=-1, keepdims=True)).astype(dt)
        angle = rng.uniform(0, 2 * np.pi, n + (1,)).astype(dt)
        d[f"{tag}/axis"], d[f"{tag}/angle"] = axis, angle
        q = ref_q.from_angle_axis(angle, axis)
        d[f"{tag}/from_angle_axis"] = q
        sa = (axis * angle).astype(dt)
        d[f"{tag}/scaled"] = sa
        d[f"{tag}/from_scaled_angle_axis"] = ref_q.from_scaled_angle_axis(sa)
        ta, tx = ref_q.to_angle_axis(q)
        d[f"{tag}/to_angle_axis_angle"], d[f"{tag}/to_angle_axis_axis"] = ta, tx
        d[f"{tag}/to_scaled_angle_axis"] = ref_q.to_scaled_angle_axis(q)
        qid = np.array([[1, 0, 0, 0], [-1, 0, 0, 0], [1, 1e-9, 0, 0]], dtype=dt)  # s <= 1e-8: zero axis (quat.py:268-271)
        ia, ix = ref_q.to_angle_axis(qid)
        d[f"{tag}/qid"], d[f"{tag}/qid_angle"], d[f"{tag}/qid_axis"] = qid, ia, ix
        euler = rng.uniform(0, 2 * np.pi, n + (3,)).astype(dt)
        order = orders6[rng.integers(0, 6, n)]
        d[f"{tag}/euler"], d[f"{tag}/order"] = euler, order
        qe = ref_q.from_euler(euler, order)
        d[f"{tag}/from_euler"] = qe
        d[f"{tag}/to_euler"] = ref_q.to_euler(qe.astype(dt), order)
        one = np.broadcast_to(np.array(["z", "x", "y"]), n + (3,))  # the BVH channel order, one order for the array
        d[f"{tag}/from_euler_zxy"] = ref_q.from_euler(euler, one)
        d[f"{tag}/to_euler_zxy"] = ref_q.to_euler(qe.astype(dt), one)
        # unroll: random signs over 96 frames x 7 joints, one exact-zero entry (dot == 0 keeps the sign)
        base = ref_q.normalize(rng.standard_normal((1, 7, 4)) + 0.15 * np.cumsum(rng.standard_normal((96, 7, 4)), axis=0))
        flips = rng.choice([-1.0, 1.0], size=(96, 7, 1))
        uq = (base * flips).astype(dt)
        uq[50, 3] = 0
        d[f"{tag}/unroll_in"] = uq
        d[f"{tag}/unroll_axis0"] = ref_q.unroll(uq.copy(), 0)
        uq1 = np.ascontiguousarray(uq.swapaxes(0, 1))
        d[f"{tag}/unroll_axis1"] = ref_q.unroll(uq1.copy(), 1)
        # slerp
        qa = ref_q.normalize(rng.standard_normal(n + (4,))).astype(dt)
        qb = ref_q.normalize(rng.standard_normal(n + (4,))).astype(dt)
        qb[0, :3] = qa[0, :3]       # identical ends: theta = 0
        qb[0, 3:6] = -qa[0, 3:6]    # opposite ends
        t = rng.uniform(-0.2, 1.2, n + (1,)).astype(dt)
        d[f"{tag}/slerp_q0"], d[f"{tag}/slerp_q1"], d[f"{tag}/slerp_t"] = qa, qb, t
        d[f"{tag}/slerp"] = ref_q.slerp(qa, qb, t)
        d[f"{tag}/slerp_long"] = ref_q.slerp(qa, qb, t, shortest=False)
        d[f"{tag}/slerp_scalar"] = ref_q.slerp(qa, qb, 0.3)
        # from_to / from_to_axis: random, parallel, anti-parallel (along x and not), one-dimensional
        v1 = rng.standard_normal((64, 3)).astype(dt)
        v2 = rng.standard_normal((64, 3)).astype(dt)
        v2[:4] = v1[:4] * 2.5
        v2[4:8] = -v1[4:8] * 0.5
        v1[8], v2[8] = [1, 0, 0], [-3, 0, 0]
        v1[9], v2[9] = [-2, 0, 0], [1, 0, 0]
        d[f"{tag}/v1"], d[f"{tag}/v2"] = v1, v2
        with np.errstate(invalid="ignore"):
            d[f"{tag}/from_to"] = ref_q.from_to(v1, v2)
            d[f"{tag}/from_to_raw"] = ref_q.from_to(ref_q.normalize(v1).astype(dt), ref_q.normalize(v2).astype(dt), normalize_input=False)
            d[f"{tag}/from_to_1d"] = ref_q.from_to(v1[20], v2[20])
            ax = ref_q.normalize(rng.standard_normal((64, 3))).astype(dt)
            d[f"{tag}/ft_axis"] = ax
            d[f"{tag}/from_to_axis"] = ref_q.from_to_axis(v1, v2, ax)
            d[f"{tag}/from_to_axis_1d"] = ref_q.from_to_axis(v1[20], v2[20], ax[20])
        # dual quaternions
        dq = rng.standard_normal(n + (8,)).astype(dt)
        d[f"{tag}/dq_raw"] = dq
        d[f"{tag}/dq_normalize_raw"] = ref_dq.normalize(dq)          # not unit after scaling: ortho branch (dual_quat.py:105-111)
        unit = ref_dq.from_rotation_translation(qa, rng.standard_normal(n + (3,)).astype(dt)) * dt(3.0)
        d[f"{tag}/dq_scaled_unit"] = unit.astype(dt)
        d[f"{tag}/dq_normalize_unit"] = ref_dq.normalize(unit.astype(dt))  # unit after scaling: no projection
        d[f"{tag}/dq_is_unit"] = np.array([ref_dq.is_unit(dq), ref_dq.is_unit(ref_dq.normalize(unit.astype(dt))),
                                           ref_dq.is_unit(np.zeros((4, 8), dtype=dt))])
        udq = np.concatenate([uq, rng.standard_normal((96, 7, 4)).astype(dt)], axis=-1)
        d[f"{tag}/dq_unroll_in"] = udq
        d[f"{tag}/dq_unroll_axis0"] = ref_dq.unroll(udq.copy(), 0)
    np.savez_compressed(os.path.join(OUT, "quat_ext.npz"), **d)
    return d


def gen_ik():
    """from_root_positions and the three mirror modes (real reference outputs)."""
    d = {}
    mapping22 = np.array([0, 5, 6, 7, 8, 1, 2, 3, 4, 9, 10, 11, 12, 13, 18, 19, 20, 21, 14, 15, 16, 17])  # legs / arms swapped
    d["body22/joints_mapping"] = mapping22
    rng = np.random.default_rng(99)
    d["end_sites"] = rng.standard_normal((5, 3)).astype(np.float32)
    for name, frames in (("chain3", 8), ("body22", 32), ("smplh52", 12), ("deep65", 10)):
        par = np.array(TOPOLOGIES[name])
        rot, gpos, off = synth_numpy(frames, par, seed=17 + len(par))
        d[f"{name}/parents"], d[f"{name}/rot"], d[f"{name}/gpos"], d[f"{name}/offsets"] = par, rot, gpos, off
        pos, _ = ref_sk.fk(rot, gpos, off, par)
        centred = (pos - pos[:, 0:1]).astype(np.float32)
        d[f"{name}/centred"] = centred
        with np.errstate(invalid="ignore"):
            d[f"{name}/from_root_positions"] = ref_sk.from_root_positions(centred, par, off)
            for axis in ("XYZ" if name == "body22" else "Y"):
                out = ref_sk.mirror(rot.copy(), gpos.copy(), par, off.copy(), d["end_sites"].copy(), None, "all", axis)
                for key, val in zip(("rot", "gpos", "offsets", "ends"), out):
                    d[f"{name}/mirror_all_{axis}/{key}"] = val
            out = ref_sk.mirror(rot.copy(), gpos.copy(), par, off.copy(), None, None, "positions", "X")
            d[f"{name}/mirror_positions_X/rot"], d[f"{name}/mirror_positions_X/gpos"] = out[0], out[1]
            if name == "body22":
                for axis in "XZ":
                    out = ref_sk.mirror(rot.copy(), gpos.copy(), par, off.copy(), None, mapping22, "symmetry", axis)
                    d[f"{name}/mirror_symmetry_{axis}/rot"], d[f"{name}/mirror_symmetry_{axis}/gpos"] = out[0], out[1]
    np.savez_compressed(os.path.join(OUT, "ik.npz"), **d)
    return d


def gen_misc():
    """SURVEY 8f rank 3 (ortho6d) and rank 4 (center_of_mass, interpolate_positions, vector.normalize)."""
    d = {}
    rng = np.random.default_rng(2024)
    for tag, dt in (("f32", np.float32), ("f64", np.float64)):
        q = ref_q.normalize(rng.standard_normal((3, 17, 4))).astype(dt)
        m = ref_q.to_matrix(q).astype(dt)
        d[f"{tag}/q"], d[f"{tag}/m"] = q, m
        d[f"{tag}/o6_from_quat"] = ref_o6.from_quat(q)
        d[f"{tag}/o6_from_matrix"] = np.ascontiguousarray(ref_o6.from_matrix(m))
        o6 = (ref_o6.from_matrix(m) * rng.uniform(0.5, 2.0, (3, 17, 1, 2)) + 0.05 * rng.standard_normal((3, 17, 3, 2))).astype(dt)
        d[f"{tag}/o6"] = o6                       # scaled and skewed: Gram-Schmidt has work to do
        d[f"{tag}/o6_to_matrix"] = np.ascontiguousarray(ref_o6.to_matrix(o6))
        d[f"{tag}/o6_to_quat"] = ref_o6.to_quat(o6)
        joints = rng.standard_normal((4, 5, 22, 3)).astype(dt)
        w = rng.uniform(0.1, 1.0, 22)
        w /= w.sum()
        d[f"{tag}/joints"], d[f"{tag}/weights"] = joints, w.astype(dt)
        d[f"{tag}/com"] = ref_com.center_of_mass(joints, w.astype(dt))
        wpf = rng.uniform(0.1, 1.0, (4, 5, 22)).astype(dt)
        d[f"{tag}/weights_pf"] = wpf
        d[f"{tag}/com_pf"] = ref_com.center_of_mass(joints, wpf)
        parts = [joints[..., 0:6, :], joints[..., 6:10, :], joints[..., 10:14, :], joints[..., 14:18, :], joints[..., 18:22, :]]
        d[f"{tag}/human_com"] = ref_com.human_center_of_mass(*parts)
        t0 = np.sort(rng.uniform(0, 10, 40))
        ts = np.concatenate([rng.uniform(-1, 11, 60), t0[[0, 7, 39]]])  # outside the range and exactly on knots
        d[f"{tag}/t_orig"], d[f"{tag}/t_sample"] = t0, ts
        p1 = rng.standard_normal((40, 3)).astype(dt)
        p2 = rng.standard_normal((4, 40, 3)).astype(dt)
        d[f"{tag}/interp_p1"], d[f"{tag}/interp_p2"] = p1, p2
        d[f"{tag}/interp_1"] = ref_time.interpolate_positions(ts, t0, p1, 0)
        d[f"{tag}/interp_2"] = ref_time.interpolate_positions(ts, t0, p2, 1)
        v = rng.standard_normal((6, 11, 3)).astype(dt)
        v[0, 0] = 0
        d[f"{tag}/vec"] = v
        d[f"{tag}/vec_normalize"] = ref_vec.normalize(v)
        d[f"{tag}/vec_normalize_eps"] = ref_vec.normalize(v, eps=1e-3)
        v5 = rng.standard_normal((9, 5)).astype(dt)
        d[f"{tag}/vec5"], d[f"{tag}/vec5_normalize"] = v5, ref_vec.normalize(v5)
    # hand-written sample of ops/tests/test_center_of_mass.py:15-40
    d["hand/com_joints"] = np.array([[[0, 0, 0], [1, 1, 1], [2, 2, 2]], [[1, 0, 0], [0, 1, 0], [0, 0, 1]]], dtype=np.float64)
    d["hand/com_weights"] = np.array([0.2, 0.3, 0.5])
    d["hand/com"] = ref_com.center_of_mass(d["hand/com_joints"], d["hand/com_weights"])
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **d)
    return d


def gen_bvh():
    """the numeric chain of BVH.get_data (io/bvh.py:332-365) run through the REAL BVH class on a
    synthetic, in-memory `data` dictionary (no file is parsed: text I/O is out of scope).  Euler angles in degrees
    with large frame-to-frame jumps so that the unroll has covers to fix, mixed per-joint channel orders."""
    import pymotion.io.bvh as ref_bvh

    d = {}
    rng = np.random.default_rng(77)
    orders = [list(o) for o in ("zyx", "xyz", "yzx", "zxy", "xzy", "yxz")]
    for tag, n_frames, n_joints in (("small", 7, 3), ("body22", 300, 22), ("long", 1000, 5)):
        # a smooth motion plus jumps of whole turns: the same rotations, the other quaternion cover
        base = np.cumsum(rng.normal(0, 6.0, (n_frames, n_joints, 3)), axis=0) + rng.uniform(-180, 180, (1, n_joints, 3))
        turns = 360.0 * rng.integers(-1, 2, (n_frames, n_joints, 3)) * (rng.uniform(size=(n_frames, n_joints, 3)) < 0.15)
        rotations = base + turns
        rot_order = np.array([orders[k % len(orders)] for k in range(n_joints)])
        bvh = ref_bvh.BVH()
        bvh.data = {
            "names": np.array([f"j{k}" for k in range(n_joints)]),
            "offsets": rng.normal(0, 0.15, (n_joints, 3)),
            "end_sites": np.zeros((0, 3)),
            "end_sites_parents": np.zeros((0,), dtype=int),
            "parents": np.array([0] + [max(0, k - 1) for k in range(1, n_joints)]),
            "rot_order": rot_order,
            "positions": np.zeros((n_frames, n_joints, 3)),
            "rotations": rotations,
            "frame_time": 1.0 / 60.0,
        }
        rots, pos, parents, offsets, end_sites, end_sites_parents = bvh.get_data()
        d[f"{tag}/rotations_deg"] = rotations
        d[f"{tag}/rot_order"] = rot_order
        d[f"{tag}/rots"] = rots
    np.savez_compressed(os.path.join(OUT, "bvh.npz"), **d)
    return d


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for fn in (gen_fk, gen_dq, gen_quat, gen_quat_ext, gen_ik, gen_misc, gen_bvh):
        if only and fn.__name__ not in only:
            continue
        out = fn()
        print(fn.__name__, len(out), "arrays")
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)}
    print(sizes)
