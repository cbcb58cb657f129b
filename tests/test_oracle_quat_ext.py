"""Pins the oracle's restatement of the REST of the quat / dual_quat surface against
fixtures written by the real reference (oracle/gen_golden.py::gen_quat_ext) and the hand-written samples of
rotations/tests/test_quat.py.  CPU only.  Same arithmetic in the same dtype => bit-exact (rtol = atol = 0)."""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import pymotion_oracle as orc

TAGS = ("f32", "f64")


def same(got, want):
    assert got.dtype == want.dtype and got.shape == want.shape
    assert_array_equal(got, want)


@pytest.mark.parametrize("tag", TAGS)
def test_angle_axis_family(golden_quat_ext, tag):
    g = golden_quat_ext
    same(orc.quat_from_angle_axis(g[f"{tag}/angle"], g[f"{tag}/axis"]), g[f"{tag}/from_angle_axis"])
    same(orc.quat_from_scaled_angle_axis(g[f"{tag}/scaled"]), g[f"{tag}/from_scaled_angle_axis"])
    q = g[f"{tag}/from_angle_axis"]
    angle, axis = orc.quat_to_angle_axis(q)
    same(angle, g[f"{tag}/to_angle_axis_angle"])
    same(axis, g[f"{tag}/to_angle_axis_axis"])
    same(orc.quat_to_scaled_angle_axis(q), g[f"{tag}/to_scaled_angle_axis"])
    angle, axis = orc.quat_to_angle_axis(g[f"{tag}/qid"])  # near-identity: zero axis
    same(angle, g[f"{tag}/qid_angle"])
    same(axis, g[f"{tag}/qid_axis"])


def test_hand_written_samples(golden_quat_ext):
    g = golden_quat_ext
    atol = 1e-6  # rotations/tests/test_quat.py atol
    assert_allclose(orc.quat_from_angle_axis(g["hand/aa_angle"], g["hand/aa_axis"]), g["hand/aa_quat"], atol=atol)
    angle, axis = orc.quat_to_angle_axis(g["hand/aa_quat"])
    assert_allclose(angle, g["hand/aa_angle"], atol=atol)
    assert_allclose(axis, g["hand/aa_axis"], atol=atol)
    assert_allclose(orc.quat_from_scaled_angle_axis(g["hand/saa"]), g["hand/saa_quat"], atol=atol)
    assert_allclose(orc.quat_to_scaled_angle_axis(g["hand/saa_quat"]), g["hand/saa"], atol=atol)
    assert_allclose(orc.quat_from_euler(g["hand/euler"], g["hand/euler_order"]), g["hand/euler_quat"], atol=atol)
    assert_allclose(orc.quat_to_euler(g["hand/euler_quat"], g["hand/euler_order"]), g["hand/euler"], atol=atol)
    q1, q2 = g["hand/slerp_q1"], g["hand/slerp_q2"]
    assert_allclose(orc.quat_slerp(q1, q2, g["hand/slerp_t"]), g["hand/slerp_gt"], atol=atol)
    assert_allclose(orc.quat_slerp(q1, q2, g["hand/slerp_t2"]), g["hand/slerp_gt2"], atol=atol)
    assert_allclose(orc.quat_slerp(q1, q2, 0.75), g["hand/slerp_gt3"], atol=atol)
    # test_quat.py:494-505: antipodal ends
    q = orc.quat_from_angle_axis(np.array([np.pi / 2]), np.array([0, 1, 1]) / np.sqrt(2))
    assert_allclose(orc.quat_slerp(q, -q, 0.5), q, atol=1e-3)
    assert_allclose(orc.quat_slerp(q, -q, 0.25, shortest=False), [0.5, 0.0, 0.353553, 0.353553], atol=1e-3)


@pytest.mark.parametrize("tag", TAGS)
def test_euler(golden_quat_ext, tag):
    g = golden_quat_ext
    same(orc.quat_from_euler(g[f"{tag}/euler"], g[f"{tag}/order"]), g[f"{tag}/from_euler"])
    dt = g[f"{tag}/euler"].dtype
    same(orc.quat_to_euler(g[f"{tag}/from_euler"].astype(dt), g[f"{tag}/order"]), g[f"{tag}/to_euler"])
    one = np.broadcast_to(np.array(["z", "x", "y"]), g[f"{tag}/euler"].shape)
    same(orc.quat_from_euler(g[f"{tag}/euler"], one), g[f"{tag}/from_euler_zxy"])
    same(orc.quat_to_euler(g[f"{tag}/from_euler"].astype(dt), one), g[f"{tag}/to_euler_zxy"])
    with pytest.raises(KeyError):
        orc.quat_from_euler(g[f"{tag}/euler"][:1, :1], np.array([[["x", "y", "w"]]]))


@pytest.mark.parametrize("tag", TAGS)
def test_unroll(golden_quat_ext, tag):
    g = golden_quat_ext
    q = g[f"{tag}/unroll_in"]
    keep = q.copy()
    same(orc.quat_unroll(q, 0), g[f"{tag}/unroll_axis0"])
    assert_array_equal(q, keep)  # the restatement does not flip in place (the reference does)
    same(orc.quat_unroll(np.ascontiguousarray(q.swapaxes(0, 1)), 1), g[f"{tag}/unroll_axis1"])
    same(orc.dq_unroll(g[f"{tag}/dq_unroll_in"], 0), g[f"{tag}/dq_unroll_axis0"])
    out = g[f"{tag}/unroll_axis0"]
    assert (np.sum(out[1:] * out[:-1], axis=-1) >= 0).all()


@pytest.mark.parametrize("tag", TAGS)
def test_slerp(golden_quat_ext, tag):
    g = golden_quat_ext
    q0, q1, t = g[f"{tag}/slerp_q0"], g[f"{tag}/slerp_q1"], g[f"{tag}/slerp_t"]
    same(orc.quat_slerp(q0, q1, t), g[f"{tag}/slerp"])
    same(orc.quat_slerp(q0, q1, t, shortest=False), g[f"{tag}/slerp_long"])
    same(orc.quat_slerp(q0, q1, 0.3), g[f"{tag}/slerp_scalar"])


@pytest.mark.parametrize("tag", TAGS)
def test_from_to(golden_quat_ext, tag):
    g = golden_quat_ext
    v1, v2, ax = g[f"{tag}/v1"], g[f"{tag}/v2"], g[f"{tag}/ft_axis"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)  # sqrt of a rounding-negative 1 - dot, overwritten by the parallel case
        same(orc.quat_from_to(v1, v2), g[f"{tag}/from_to"])
        n1, n2 = orc.quat_normalize(v1).astype(v1.dtype), orc.quat_normalize(v2).astype(v1.dtype)
        same(orc.quat_from_to(n1, n2, normalize_input=False), g[f"{tag}/from_to_raw"])
        same(orc.quat_from_to(v1[20], v2[20]), g[f"{tag}/from_to_1d"])
        same(orc.quat_from_to_axis(v1, v2, ax), g[f"{tag}/from_to_axis"])
        same(orc.quat_from_to_axis(v1[20], v2[20], ax[20]), g[f"{tag}/from_to_axis_1d"])
    rot = g[f"{tag}/from_to"]
    assert_array_equal(rot[:4], np.tile([1.0, 0, 0, 0], (4, 1)))  # parallel -> identity
    assert_array_equal(rot[4:10, 0], 0.0)                          # anti-parallel -> half turn
    # the reference's own property (test_quat.py:571-584): the rotation takes v1 onto v2
    moved = orc.quat_mul_vec(rot[10:], v1[10:])
    assert_allclose(moved / np.linalg.norm(moved, axis=-1, keepdims=True),
                    v2[10:] / np.linalg.norm(v2[10:], axis=-1, keepdims=True), atol=1e-3)


@pytest.mark.parametrize("tag", TAGS)
def test_dual_quat_normalize_is_unit(golden_quat_ext, tag):
    g = golden_quat_ext
    same(orc.dq_normalize(g[f"{tag}/dq_raw"]), g[f"{tag}/dq_normalize_raw"])
    same(orc.dq_normalize(g[f"{tag}/dq_scaled_unit"]), g[f"{tag}/dq_normalize_unit"])
    flags = [orc.dq_is_unit(g[f"{tag}/dq_raw"]), orc.dq_is_unit(g[f"{tag}/dq_normalize_unit"]),
             orc.dq_is_unit(np.zeros((4, 8), dtype=g[f"{tag}/dq_raw"].dtype))]
    assert flags == g[f"{tag}/dq_is_unit"].tolist() == [False, True, True]
    assert orc.dq_is_unit(g[f"{tag}/dq_normalize_raw"])  # test_dual_quat.py:51-75
