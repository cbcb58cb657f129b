"""Pins oracle/ (NumPy and C restatements) against the fixtures written by the
real reference (oracle/gen_golden.py) and the reference tests' hand-written
golden arrays.  CPU only."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import oracle_c
from oracle import pymotion_oracle as orc

SKELS = ("body22", "smplh52", "deep65")
TIGHT = dict(rtol=1e-13, atol=1e-13)  # float64 outputs: same arithmetic, only BLAS summation order may differ


# ---------------------------------------------------------------- fk
@pytest.mark.parametrize("case", ["chain3_ident", "chain3_rot"])
def test_fk_chain3_matches_reference_and_hand_goldens(golden_fk, case):
    g = golden_fk
    pos, rotm = orc.fk(g[f"{case}/rot"], g["chain3/gpos"], g["chain3/offsets"], g["chain3/parents"])
    assert_allclose(pos, g[f"{case}/pos"], **TIGHT)
    assert_allclose(rotm, g[f"{case}/rotm"], **TIGHT)
    # hand-written goldens of ops/tests/test_skeleton.py at its own atol (1e-6)
    assert_allclose(pos, g[f"{case}/hand_pos"], atol=1e-6)
    if case == "chain3_rot":
        assert_allclose(rotm, g["chain3_rot/hand_rotm"], atol=1e-6)
    else:
        assert_allclose(rotm, np.tile(np.eye(3), (2, 3, 1, 1)), atol=1e-6)
        pos_pf, rotm_pf = orc.fk(g[f"{case}/rot"], g["chain3/gpos"], np.tile(g["chain3/offsets"], (2, 1, 1)), g["chain3/parents"])
        assert_allclose(pos_pf, g["chain3_ident/pos_pf"], **TIGHT)
        assert_allclose(rotm_pf, g["chain3_ident/rotm_pf"], **TIGHT)


def test_fk_nd_leading_dims(golden_fk):
    g = golden_fk
    pos, rotm = orc.fk(g["chain3_nd/rot"], g["chain3_nd/gpos"], g["chain3/offsets"], g["chain3/parents"])
    assert pos.shape == (4, 3, 4, 3, 3) and rotm.shape == (4, 3, 4, 3, 3, 3)
    assert_allclose(pos, g["chain3_nd/pos"], **TIGHT)
    assert_allclose(rotm, g["chain3_nd/rotm"], **TIGHT)


@pytest.mark.parametrize("name", SKELS)
def test_fk_skeletons(golden_fk, name):
    g = golden_fk
    par, rot, gp, off = (g[f"{name}/{k}"] for k in ("parents", "rot", "gpos", "offsets"))
    pos, rotm = orc.fk(rot, gp, off, par)
    assert pos.dtype == np.float64 and rotm.dtype == np.float64  # reference promotes (skeleton.py:44)
    assert_allclose(pos, g[f"{name}/pos"], **TIGHT)
    assert_allclose(rotm, g[f"{name}/rotm"], **TIGHT)
    pos, rotm = orc.fk(rot, gp, g[f"{name}/offsets_pf"], par)
    assert_allclose(pos, g[f"{name}/pos_pf"], **TIGHT)
    assert_allclose(rotm, g[f"{name}/rotm_pf"], **TIGHT)
    pos, rotm = orc.fk(rot.astype(np.float64), gp.astype(np.float64), off.astype(np.float64), par)
    assert_allclose(pos, g[f"{name}/pos_f64"], **TIGHT)
    assert_allclose(rotm, g[f"{name}/rotm_f64"], **TIGHT)
    # C restatement: float32 locals, float64 chain
    cpos, crotm = oracle_c.fk(rot, gp, off, par)
    assert_allclose(cpos, g[f"{name}/pos"], rtol=1e-6, atol=1e-6)
    assert_allclose(crotm, g[f"{name}/rotm"], rtol=1e-6, atol=1e-6)
    cpos, crotm = oracle_c.fk(rot, gp, g[f"{name}/offsets_pf"], par)
    assert_allclose(cpos, g[f"{name}/pos_pf"], rtol=1e-6, atol=1e-6)
    assert_allclose(crotm, g[f"{name}/rotm_pf"], rtol=1e-6, atol=1e-6)


def test_fk_edge_shapes(golden_fk):
    g = golden_fk
    pos, rotm = orc.fk(g["unbatched/rot"], g["unbatched/gpos"], g["unbatched/offsets"], g["unbatched/parents"])
    assert pos.shape == (22, 3) and rotm.shape == (22, 3, 3)
    assert_allclose(pos, g["unbatched/pos"], **TIGHT)
    assert_allclose(rotm, g["unbatched/rotm"], **TIGHT)
    pos, rotm = orc.fk(g["single/rot"], g["single/gpos"], g["single/offsets"], np.array([0]))
    assert_allclose(pos, g["single/pos"], **TIGHT)
    assert_allclose(rotm, g["single/rotm"], **TIGHT)
    pos, rotm = orc.fk(g["bcast/rot"], np.zeros((1, 3)), g["bcast/offsets"], g["body22/parents"])
    assert_allclose(pos, g["bcast/pos"], **TIGHT)
    assert_allclose(rotm, g["bcast/rotm"], **TIGHT)
    cpos, crotm = oracle_c.fk(g["bcast/rot"], np.zeros((1, 3)), g["bcast/offsets"], g["body22/parents"])
    assert_allclose(cpos, g["bcast/pos"], rtol=1e-6, atol=1e-6)
    assert_allclose(crotm, g["bcast/rotm"], rtol=1e-6, atol=1e-6)


def test_fk_zero_quaternion_is_identity(golden_fk):
    g = golden_fk
    # rot[0,1] was zeroed by the generator: q/(0+1e-8) = 0 -> R = I (SURVEY 8a notes)
    assert_array_equal(g["body22/rot"][0, 1], 0)
    _, rotm = orc.fk(g["body22/rot"][:1], g["body22/gpos"][:1], g["body22/offsets"], g["body22/parents"])
    assert_allclose(rotm[0, 1], rotm[0, 0], atol=0)  # joint 1 hangs off the root with an identity local rotation


# ---------------------------------------------------------------- dual quaternions
@pytest.mark.parametrize("case", ["chain3_ident", "chain3_rot"])
def test_dq_chain3(golden_dq, case):
    g = golden_dq
    par, off, gp = g["chain3/parents"], g["chain3/offsets"], g["chain3/gpos"]
    dq = orc.to_root_dual_quat(g[f"{case}/rot"], gp, par, off)
    assert_allclose(dq, g[f"{case}/dq"], **TIGHT)
    rr, tt = orc.dq_to_rotation_translation(dq)
    assert_allclose(rr, g[f"{case}/root_rot"], **TIGHT)
    assert_allclose(tt, g[f"{case}/root_trans"], **TIGHT)
    # hand-written goldens (test_skeleton.py:42-59, :123-171)
    assert_allclose(tt[:, 1:], g[f"{case}/hand_root_trans"][:, 1:], atol=1e-6)
    assert_allclose(tt[:, 0], gp, atol=1e-6)
    if case == "chain3_rot":
        assert_allclose(rr, orc.quat_from_matrix(g["chain3_rot/hand_root_rotm"]), atol=1e-6)
    trans, rots = orc.from_root_dual_quat(dq, par)
    assert_allclose(trans, g[f"{case}/back_trans"], **TIGHT)
    assert_allclose(rots, g[f"{case}/back_rot"], **TIGHT)
    # the inverse recovers the inputs (test_skeleton.py:70-77, :183-190)
    assert_allclose(rots, g[f"{case}/rot"], atol=1e-6)
    assert_allclose(trans[:, 1:], np.tile(off[1:], (2, 1, 1)), atol=1e-6)
    assert_allclose(trans[:, 0], gp, atol=1e-6)


@pytest.mark.parametrize("name", SKELS)
def test_dq_skeletons(golden_dq, name):
    g = golden_dq
    par, rot, gp, off = (g[f"{name}/{k}"] for k in ("parents", "rot", "gpos", "offsets"))
    dq = orc.to_root_dual_quat(rot, gp, par, off)
    assert dq.dtype == np.float64
    assert_allclose(dq, g[f"{name}/dq"], **TIGHT)
    trans, rots = orc.from_root_dual_quat(dq, par)
    assert_allclose(trans, g[f"{name}/back_trans"], **TIGHT)
    assert_allclose(rots, g[f"{name}/back_rot"], **TIGHT)
    t32, r32 = orc.from_root_dual_quat(g[f"{name}/dq"].astype(np.float32), par)
    assert t32.dtype == np.float32
    assert_allclose(t32, g[f"{name}/back_trans_f32in"], rtol=0, atol=0)
    assert_allclose(r32, g[f"{name}/back_rot_f32in"], rtol=0, atol=0)
    # C restatement
    cdq = oracle_c.to_root_dual_quat(rot, gp, par, off)
    assert_allclose(cdq, g[f"{name}/dq"], rtol=1e-6, atol=1e-6)
    ct, cr = oracle_c.from_root_dual_quat(g[f"{name}/dq"], par)
    assert_allclose(ct, g[f"{name}/back_trans"], rtol=1e-12, atol=1e-12)
    assert_allclose(cr, g[f"{name}/back_rot"], rtol=1e-12, atol=1e-12)


def test_to_root_dual_quat_asserts_on_root_offset(golden_dq):
    g = golden_dq
    off = g["body22/offsets"].copy()
    off[0, 1] = 0.5
    with pytest.raises(AssertionError):
        orc.to_root_dual_quat(g["body22/rot"], g["body22/gpos"], g["body22/parents"], off)
    with pytest.raises(AssertionError):
        oracle_c.to_root_dual_quat(g["body22/rot"], g["body22/gpos"], g["body22/parents"], off)


def test_from_global_rotations(golden_dq):
    g = golden_dq
    assert_allclose(orc.from_global_rotations(g["fgr/global"], g["fgr/parents"]), g["fgr/local"], rtol=0, atol=0)


# ---------------------------------------------------------------- quaternion / dual-quaternion primitives
def test_quat_hand_goldens(golden_quat):
    g = golden_quat
    assert_allclose(orc.quat_mul(g["hand/qa"], g["hand/qb"]), g["hand/mul_ab"], atol=1e-6)
    assert_allclose(orc.quat_mul(g["hand/qb"], g["hand/qa"]), g["hand/mul_ba"], atol=1e-6)
    rotated = orc.quat_mul_vec(g["hand/qa"], g["hand/v"])
    assert_allclose(rotated, g["hand/mul_vec"], atol=1e-6)
    assert_allclose(orc.quat_mul_vec(orc.quat_inverse(g["hand/qa"]), rotated), g["hand/v"], atol=1e-6)
    assert_allclose(orc.quat_to_matrix(g["hand/qa"]), g["hand/matrix"], atol=1e-6)
    assert_allclose(orc.quat_from_matrix(g["hand/matrix"]), g["hand/qa"], atol=1e-6)


@pytest.mark.parametrize("tag", ["f32", "f64"])
def test_quat_primitives_bit_exact(golden_quat, tag):
    g = golden_quat
    q0, q1, v, qu, t = (g[f"{tag}/{k}"] for k in ("q0", "q1", "v", "qu", "t"))
    exact = dict(rtol=0, atol=0)
    checks = {
        "mul": orc.quat_mul(q0, q1),
        "mul_bcast": orc.quat_mul(q0[:, :1], q1),
        "mul_vec": orc.quat_mul_vec(q0, v),
        "length": orc.quat_length(q0),
        "normalize": orc.quat_normalize(q0),
        "normalize_eps": orc.quat_normalize(q0, eps=1e-2),
        "conjugate": orc.quat_conjugate(q0),
        "inverse": orc.quat_inverse(q0),
        "to_matrix": orc.quat_to_matrix(q0),
        "unit_matrix": orc.quat_to_matrix(qu),
        "from_matrix": orc.quat_from_matrix(g[f"{tag}/unit_matrix"].astype(q0.dtype)),
        "dq": orc.dq_from_rotation_translation(qu, t),
        "dq_from_translation": orc.dq_from_translation(t),
    }
    for key, got in checks.items():
        want = g[f"{tag}/{key}"]
        assert got.dtype == want.dtype, key
        assert_allclose(got, want, err_msg=key, **exact)
    rr, tt = orc.dq_to_rotation_translation(g[f"{tag}/dq"])
    assert_allclose(rr, g[f"{tag}/dq_rot"], **exact)
    assert_allclose(tt, g[f"{tag}/dq_trans"], **exact)
    # round trip of the reference test (test_dual_quat.py:14-49) at its 1e-6
    assert_allclose(tt, t, atol=1e-6)


def test_from_matrix_branches(golden_quat):
    g = golden_quat
    m = g["branches/matrix"]
    # make sure the fixture really exercises all four branches of quat.py:111-155
    taken = {(bool(a[2, 2] < 0), bool(a[0, 0] > a[1, 1]), bool(a[0, 0] < -a[1, 1])) for a in m}
    assert len({(n, x if n else z) for n, x, z in taken}) == 4
    assert_allclose(orc.quat_from_matrix(m), g["branches/quat"], rtol=0, atol=0)
