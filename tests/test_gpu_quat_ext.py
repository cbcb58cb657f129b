"""Parity of the CUDA kernels for the REST of the quat / dual_quat surface against the
fixtures written by the real reference (tests/golden/quat_ext.npz) and against the oracle on larger seeded
inputs.  Needs a B200: -m gpu.  Tolerance 1e-5 unless a comment says why not."""
import warnings

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose, assert_array_equal

from oracle import pymotion_oracle as orc

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-5, atol=1e-5)
TWO_PI = 2 * np.pi


@pytest.fixture(scope="module")
def quat():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pymotion_b200.rotations.quat as mod

    return mod


@pytest.fixture(scope="module")
def dquat(quat):
    import pymotion_b200.rotations.dual_quat as mod

    return mod


def angle_diff(a, b):
    """|a - b| on the circle."""
    return np.abs((np.asarray(a, dtype=np.float64) - b + np.pi) % TWO_PI - np.pi)


def same_rotation(q, want, atol=1e-5):
    assert_allclose(np.abs(np.sum(np.asarray(q, dtype=np.float64) * want, axis=-1)), 1.0, atol=atol)


def test_angle_axis_family(quat, golden_quat_ext):
    g = golden_quat_ext
    q = quat.from_angle_axis(g["f32/angle"], g["f32/axis"])
    assert isinstance(q, np.ndarray) and q.dtype == np.float32 and q.shape == (3, 29, 4)
    assert_allclose(q, g["f32/from_angle_axis"], **TOL)
    assert_allclose(quat.from_scaled_angle_axis(g["f32/scaled"]), g["f32/from_scaled_angle_axis"], **TOL)
    angle, axis = quat.to_angle_axis(g["f32/from_angle_axis"])
    assert angle.shape == (3, 29, 1) and axis.shape == (3, 29, 3)
    assert_allclose(angle, g["f32/to_angle_axis_angle"], **TOL)
    # axis = xyz / sqrt(1 - w^2): the fp32 rounding of 1 - w*w (fused on the GPU, two roundings in NumPy) is
    # amplified by 1 / (1 - w^2) near the identity; 1e-4 there, 1e-5 elsewhere
    far = np.abs(g["f32/from_angle_axis"][..., 0]) < 0.99
    assert_allclose(axis[far], g["f32/to_angle_axis_axis"][far], **TOL)
    assert_allclose(axis, g["f32/to_angle_axis_axis"], rtol=1e-4, atol=1e-4)
    sa = quat.to_scaled_angle_axis(g["f32/from_angle_axis"])
    assert_allclose(sa[far], g["f32/to_scaled_angle_axis"][far], **TOL)
    assert_allclose(sa, g["f32/to_scaled_angle_axis"], rtol=1e-4, atol=1e-4)
    angle, axis = quat.to_angle_axis(g["f32/qid"])  # sin(angle / 2) <= 1e-8: zero axis (quat.py:268-271)
    assert_allclose(angle, g["f32/qid_angle"], **TOL)
    assert_array_equal(axis, g["f32/qid_axis"])
    # hand-written samples of rotations/tests/test_quat.py:55-77, :101-119 at its atol
    assert_allclose(quat.from_angle_axis(g["hand/aa_angle"].astype(np.float32), g["hand/aa_axis"].astype(np.float32)),
                    g["hand/aa_quat"], atol=1e-6)
    assert_allclose(quat.from_scaled_angle_axis(g["hand/saa"].astype(np.float32)), g["hand/saa_quat"], atol=1e-6)
    assert_allclose(quat.to_scaled_angle_axis(g["hand/saa_quat"].astype(np.float32)), g["hand/saa"], atol=1e-6)
    # the null vector: nan axis, like the reference (0 / 0)
    z = quat.from_scaled_angle_axis(np.zeros((2, 3), dtype=np.float32))
    assert_array_equal(z[:, 0], 1.0)
    assert np.isnan(z[:, 1:]).all()
    # torch tensors stay torch tensors on their device
    t = quat.from_angle_axis(torch.from_numpy(g["f32/angle"]).cuda(), torch.from_numpy(g["f32/axis"]).cuda())
    assert t.is_cuda and t.dtype == torch.float32
    assert_allclose(t.cpu().numpy(), g["f32/from_angle_axis"], **TOL)


def test_euler(quat, golden_quat_ext):
    g = golden_quat_ext
    euler, order = g["f32/euler"], g["f32/order"]
    q = quat.from_euler(euler, order)
    assert_allclose(q, g["f32/from_euler"], **TOL)
    assert_allclose(quat.from_euler(euler, np.broadcast_to(np.array(["z", "x", "y"]), euler.shape)), g["f32/from_euler_zxy"], **TOL)
    q32 = g["f32/from_euler"].astype(np.float32)
    e = quat.to_euler(q32, order)
    assert e.shape == euler.shape and (e >= 0).all() and (e < TWO_PI + 1e-6).all()
    assert angle_diff(e, g["f32/to_euler"]).max() < 5e-5   # atan2 of fp32 sums; compared on the circle
    assert angle_diff(quat.to_euler(q32, np.broadcast_to(np.array(["z", "x", "y"]), euler.shape)), g["f32/to_euler_zxy"]).max() < 5e-5
    same_rotation(quat.from_euler(e, order), q32, atol=2e-5)  # the angles are a valid decomposition
    # hand-written samples, test_quat.py:150-170
    assert_allclose(quat.from_euler(g["hand/euler"].astype(np.float32), g["hand/euler_order"]), g["hand/euler_quat"], atol=1e-6)
    assert angle_diff(quat.to_euler(g["hand/euler_quat"].astype(np.float32), g["hand/euler_order"]), g["hand/euler"]).max() < 1e-6
    with pytest.raises(AssertionError):
        quat.from_euler(euler, order[:2])
    with pytest.raises(KeyError):
        quat.from_euler(euler[:1, :1], np.array([[["x", "y", "w"]]]))
    # larger, all six orders, against the oracle
    rng = np.random.default_rng(5)
    orders6 = np.array([list(p) for p in ("xyz", "xzy", "yxz", "yzx", "zxy", "zyx")])
    big_e = rng.uniform(0, TWO_PI, (20_000, 3)).astype(np.float32)
    big_o = orders6[rng.integers(0, 6, 20_000)]
    assert_allclose(quat.from_euler(big_e, big_o), orc.quat_from_euler(big_e, big_o), **TOL)


@pytest.mark.parametrize("tag", ["f32"])
def test_unroll_fixtures(quat, dquat, golden_quat_ext, tag):
    g = golden_quat_ext
    x = g[f"{tag}/unroll_in"]
    keep = x.copy()
    assert_array_equal(quat.unroll(x, 0), g[f"{tag}/unroll_axis0"])
    assert_array_equal(x, keep)  # input untouched (the NumPy reference flips in place)
    assert_array_equal(quat.unroll(np.ascontiguousarray(x.swapaxes(0, 1)), 1), g[f"{tag}/unroll_axis1"])
    assert_array_equal(quat.unroll(torch.from_numpy(x), dim=0).numpy(), g[f"{tag}/unroll_axis0"])
    assert_array_equal(dquat.unroll(g[f"{tag}/dq_unroll_in"], 0), g[f"{tag}/dq_unroll_axis0"])


@pytest.mark.parametrize("n_steps,n_cols", [(1, 5), (127, 3), (128, 1), (129, 22), (5000, 22), (20_011, 7), (260, 512), (300, 700)])
def test_unroll_scan_sizes(quat, dquat, set_knobs, n_steps, n_cols):
    """Chunk boundaries of the three-kernel scan (chunks of 128 steps), exact zeros (sign reset), one column, and both
    apply kernels (one block per chunk up to 512 columns, flat beyond)."""
    rng = np.random.default_rng(n_steps + n_cols)
    base = np.cumsum(0.2 * rng.standard_normal((n_steps, n_cols, 4)), axis=0) + rng.standard_normal((1, n_cols, 4))
    x = (base * rng.choice([-1.0, 1.0], size=(n_steps, n_cols, 1))).astype(np.float32)
    if n_steps > 200:
        x[rng.integers(1, n_steps, 5), rng.integers(0, n_cols, 5)] = 0
    want = orc.quat_unroll(x, 0)
    got = quat.unroll(x, 0)
    assert_array_equal(got, want)
    assert (np.sum(got[1:] * got[:-1], axis=-1) >= 0).all()
    set_knobs({"PMB_UNROLL_CHUNK_APPLY": "0"})
    assert_array_equal(quat.unroll(x, 0), want)
    set_knobs({"PMB_UNROLL_CHUNK_APPLY": "1"})
    dq = np.concatenate([x, rng.standard_normal((n_steps, n_cols, 4)).astype(np.float32)], axis=-1)
    assert_array_equal(dquat.unroll(dq, 0), orc.dq_unroll(dq, 0))


def test_slerp(quat, golden_quat_ext):
    g = golden_quat_ext
    q0, q1, t = g["f32/slerp_q0"], g["f32/slerp_q1"], g["f32/slerp_t"]
    # rows 0..2 of the first block have identical ends and rows 3..5 opposite ends: acos is ill-conditioned at
    # |dot| = 1 (an ulp of the dot product moves the angle by 3e-4), the result is not
    for kwargs, key in (({}, "f32/slerp"), ({"shortest": False}, "f32/slerp_long")):
        got = quat.slerp(q0, q1, t, **kwargs)
        assert_allclose(got[1:], g[key][1:], **TOL)
        assert_allclose(got[0, 6:], g[key][0, 6:], **TOL)
        if not kwargs:
            assert_allclose(got[0, :6], g[key][0, :6], atol=2e-4)
    assert_allclose(quat.slerp(q0, q1, 0.3)[1:], g["f32/slerp_scalar"][1:], **TOL)
    # hand-written samples, test_quat.py:393-452
    h1, h2 = g["hand/slerp_q1"].astype(np.float32), g["hand/slerp_q2"].astype(np.float32)
    assert_allclose(quat.slerp(h1, h2, g["hand/slerp_t"].astype(np.float32)), g["hand/slerp_gt"], atol=1e-6)
    assert_allclose(quat.slerp(h1, h2, g["hand/slerp_t2"].astype(np.float32)), g["hand/slerp_gt2"], atol=1e-6)
    assert_allclose(quat.slerp(h1, h2, 0.75), g["hand/slerp_gt3"], atol=1e-6)
    assert_allclose(quat.slerp(h1[None, None], h2[None, None], g["hand/slerp_t2"].astype(np.float32)[None, None]),
                    g["hand/slerp_gt2"][None, None], atol=1e-6)
    q = orc.quat_from_angle_axis(np.array([np.pi / 2]), np.array([0, 1, 1]) / np.sqrt(2)).astype(np.float32)
    assert_allclose(quat.slerp(q, -q, 0.5), q, atol=1e-3)                                                  # :494-499
    assert_allclose(quat.slerp(q, -q, 0.25, shortest=False), [0.5, 0.0, 0.353553, 0.353553], atol=1e-3)  # :500-505
    # larger, against the oracle
    rng = np.random.default_rng(9)
    a = orc.quat_normalize(rng.standard_normal((50_000, 4))).astype(np.float32)
    b = orc.quat_normalize(rng.standard_normal((50_000, 4))).astype(np.float32)
    tt = rng.uniform(0, 1, (50_000, 1)).astype(np.float32)
    far = np.abs(np.sum(a * b, axis=-1)) < 0.999  # away from the ill-conditioned ends
    assert_allclose(quat.slerp(a, b, tt)[far], orc.quat_slerp(a, b, tt)[far], rtol=1e-5, atol=2e-5)


def test_from_to(quat, golden_quat_ext):
    g = golden_quat_ext
    v1, v2, ax = g["f32/v1"], g["f32/v2"], g["f32/ft_axis"]
    got = quat.from_to(v1, v2)
    assert_allclose(got, g["f32/from_to"], **TOL)
    assert_array_equal(got[:4], np.tile(np.float32([1, 0, 0, 0]), (4, 1)))  # parallel -> identity
    assert_array_equal(got[4:10, 0], 0.0)                                     # anti-parallel -> half turn
    n1, n2 = orc.quat_normalize(v1).astype(np.float32), orc.quat_normalize(v2).astype(np.float32)
    assert_allclose(quat.from_to(n1, n2, normalize_input=False), g["f32/from_to_raw"], **TOL)
    one = quat.from_to(v1[20], v2[20])
    assert one.shape == (4,)
    assert_allclose(one, g["f32/from_to_1d"], **TOL)
    assert_allclose(quat.from_to_axis(v1, v2, ax), g["f32/from_to_axis"], **TOL)
    assert_allclose(quat.from_to_axis(v1[20], v2[20], ax[20]), g["f32/from_to_axis_1d"], **TOL)
    with pytest.raises(AssertionError):
        quat.from_to(v1, v2[:5])
    # the reference's own property (test_quat.py:571-584): the rotation takes v1 onto v2
    rng = np.random.default_rng(0)
    a, b = rng.random((1000, 3)).astype(np.float32), rng.random((1000, 3)).astype(np.float32)
    moved = orc.quat_mul_vec(quat.from_to(a, b).astype(np.float64), a.astype(np.float64))
    assert_allclose(moved / np.linalg.norm(moved, axis=-1, keepdims=True), b / np.linalg.norm(b, axis=-1, keepdims=True), atol=1e-3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        assert_allclose(quat.from_to(a, b), orc.quat_from_to(a, b), rtol=1e-5, atol=2e-5)


def test_dual_quat_normalize_is_unit_unroll(dquat, golden_quat_ext):
    g = golden_quat_ext
    raw, scaled = g["f32/dq_raw"], g["f32/dq_scaled_unit"]
    assert_allclose(dquat.normalize(raw), g["f32/dq_normalize_raw"], **TOL)       # projection branch (dual_quat.py:105-111)
    assert_allclose(dquat.normalize(scaled), g["f32/dq_normalize_unit"], **TOL)   # already orthogonal: scaling only
    assert dquat.is_unit(raw) is False
    assert dquat.is_unit(dquat.normalize(raw)) is True                             # test_dual_quat.py:51-75
    assert dquat.is_unit(g["f32/dq_normalize_unit"].astype(np.float32)) is True
    assert dquat.is_unit(np.zeros((4, 8), dtype=np.float32)) is True               # all-zero real parts (:131-132)
    assert dquat.is_unit(raw, atol=1e9) is False                                   # norms still off
    t = dquat.normalize(torch.from_numpy(raw).cuda())
    assert t.is_cuda
    assert_allclose(t.cpu().numpy(), g["f32/dq_normalize_raw"], **TOL)
    rng = np.random.default_rng(3)
    big = rng.standard_normal((30_000, 8)).astype(np.float32)
    assert_allclose(dquat.normalize(big), orc.dq_normalize(big), rtol=1e-5, atol=2e-5)
