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

import pymotion.ops.skeleton as ref_sk  # noqa: E402
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


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for fn in (gen_fk, gen_dq, gen_quat):
        out = fn()
        print(fn.__name__, len(out), "arrays")
    sizes = {f: os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)}
    print(sizes)
