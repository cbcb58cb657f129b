"""Parity of the ortho6d / center_of_mass / interpolate_positions / vector.normalize kernels (SURVEY 8f rank 3
tail, rank 4) against fixtures written by the real reference and the oracle.  Needs a B200: -m gpu."""
import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose, assert_array_equal

from oracle import pymotion_oracle as orc

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-5, atol=1e-5)


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pymotion_b200.ops.center_of_mass as com
    import pymotion_b200.ops.time as tm
    import pymotion_b200.ops.vector as vec
    import pymotion_b200.rotations.ortho6d as o6

    return o6, com, tm, vec


def test_ortho6d(mods, golden_misc):
    o6, *_ = mods
    g = golden_misc
    got = o6.from_quat(g["f32/q"])
    assert isinstance(got, np.ndarray) and got.dtype == np.float32 and got.shape == (3, 17, 3, 2)
    assert_allclose(got, g["f32/o6_from_quat"], **TOL)
    assert_array_equal(o6.from_matrix(g["f32/m"]), g["f32/o6_from_matrix"])
    assert_allclose(o6.to_matrix(g["f32/o6"]), g["f32/o6_to_matrix"], **TOL)
    q = o6.to_quat(g["f32/o6"])
    assert_allclose(np.abs(np.sum(q * g["f32/o6_to_quat"], axis=-1)), 1.0, atol=1e-5)
    assert (np.sum(q * g["f32/o6_to_quat"], axis=-1) > 0).mean() > 0.98
    assert_allclose(o6.to_matrix(o6.from_matrix(g["f32/m"])), g["f32/m"], atol=1e-6)  # test_ortho6d.py:14
    t = o6.to_matrix(torch.from_numpy(g["f32/o6"]).cuda())
    assert t.is_cuda
    assert_allclose(t.cpu().numpy(), g["f32/o6_to_matrix"], **TOL)
    rng = np.random.default_rng(1)
    big = rng.standard_normal((50_000, 3, 2)).astype(np.float32)
    assert_allclose(o6.to_matrix(big), orc.ortho6d_to_matrix(big), rtol=1e-4, atol=1e-4)  # skinny pairs are ill-conditioned


def test_center_of_mass(mods, golden_misc):
    _, com, *_ = mods
    g = golden_misc
    j = g["f32/joints"]
    assert_allclose(com.center_of_mass(j, g["f32/weights"]), g["f32/com"], **TOL)
    assert_array_equal(com.center_of_mass(j, g["f32/weights"]), g["f32/com"])  # same order of products and sums in fp32
    assert_allclose(com.center_of_mass(j, g["f32/weights_pf"]), g["f32/com_pf"], **TOL)
    assert_allclose(com.human_center_of_mass(j[..., 0:6, :], j[..., 6:10, :], j[..., 10:14, :], j[..., 14:18, :], j[..., 18:22, :]),
                    g["f32/human_com"], **TOL)
    assert_allclose(com.center_of_mass(np.float32([[[1, 2, 3], [4, 5, 6], [7, 8, 9]]]), np.float32([[0.3, 0.3, 0.4]])),
                    [[4.3, 5.3, 6.3]], atol=1e-6)  # ops/tests/test_center_of_mass.py:30-36
    rng = np.random.default_rng(2)
    big = rng.standard_normal((100_003, 52, 3)).astype(np.float32)
    w = rng.uniform(0, 1, 52).astype(np.float32)
    assert_allclose(com.center_of_mass(big, w), orc.center_of_mass(big, w), **TOL)


def test_interpolate_positions(mods, golden_misc):
    _, _, tm, _ = mods
    g = golden_misc
    ts, t0 = g["f32/t_sample"], g["f32/t_orig"]
    got = tm.interpolate_positions(ts, t0, g["f32/interp_p1"], 0)
    assert got.shape == (63, 3)
    assert_allclose(got, g["f32/interp_1"], **TOL)
    assert_allclose(tm.interpolate_positions(ts, t0, g["f32/interp_p2"], 1), g["f32/interp_2"], **TOL)
    assert_allclose(tm.interpolate_positions(ts, t0, g["f32/interp_p2"], -2), g["f32/interp_2"], **TOL)
    # [T, J, 3] along axis 0: the case every fk caller has (the NumPy reference cannot broadcast it)
    rng = np.random.default_rng(3)
    times = np.arange(5000) / 60.0
    samples = np.sort(rng.uniform(0, times[-1], 7001))
    pos = np.cumsum(rng.standard_normal((5000, 22, 3)), axis=0).astype(np.float32)
    want = orc.interpolate_positions(samples, times, pos.astype(np.float64), 0)
    assert_allclose(tm.interpolate_positions(samples, times, pos, 0), want, rtol=1e-5, atol=1e-4)
    with pytest.raises(AssertionError):
        tm.interpolate_positions(ts, t0[:-1], g["f32/interp_p1"], 0)
    with pytest.raises(AssertionError):
        tm.interpolate_positions(ts, t0, g["f32/interp_p1"], 0, method="cubic")


def test_vector_normalize(mods, golden_misc):
    *_, vec = mods
    g = golden_misc
    assert_allclose(vec.normalize(g["f32/vec"]), g["f32/vec_normalize"], **TOL)
    assert_allclose(vec.normalize(g["f32/vec"], eps=1e-3), g["f32/vec_normalize_eps"], **TOL)
    assert_allclose(vec.normalize(g["f32/vec5"]), g["f32/vec5_normalize"], **TOL)
    assert_array_equal(vec.normalize(g["f32/vec"])[0, 0], 0.0)  # the null vector stays null (0 / eps)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 7, 1001, 4098, 65539])
def test_vector_normalize_vec3_paths(mods, set_knobs, n):
    """3-vectors go four per thread when the arrays are 16-byte aligned (the last n % 4 through the generic kernel):
    same bits as the generic kernel, on aligned and on 12-byte-offset arrays, and the oracle's values."""
    *_, vec = mods
    rng = np.random.default_rng(n)
    v = rng.normal(size=(n + 1, 3)).astype(np.float32)
    v[n // 2] = 0.0
    want = orc.vec_normalize(v)
    fast = vec.normalize(v)
    assert_allclose(fast, want, **TOL)
    set_knobs({"PMB_VEC3_X4": "0"})
    assert_array_equal(vec.normalize(v), fast)
    set_knobs({"PMB_VEC3_X4": "1"})
    t = torch.from_numpy(v).cuda()
    shifted = vec.normalize(t[1:])  # data pointer 12 bytes past a 16-byte boundary: generic kernel
    assert isinstance(shifted, torch.Tensor)
    assert_array_equal(shifted.cpu().numpy(), fast[1:])


@pytest.mark.parametrize("tag", ["small", "body22", "long"])
def test_bvh_get_data_fused(golden_bvh, tag):
    """io.bvh.rotations_to_quat / get_data: the fused from_euler -> unroll -> normalize scan against the output of the
    real BVH.get_data (fixtures) -- same rotations everywhere, and the SAME cover (sign) wherever the unroll decision
    of the float64 reference is not within fp32 rounding of a tie (consecutive frames more than ~90 degrees apart in
    quaternion space)."""
    import pymotion_b200.io.bvh as bvh
    import pymotion_b200.rotations.quat as quat

    g = golden_bvh
    rot_deg, order, want = g[f"{tag}/rotations_deg"].astype(np.float32), g[f"{tag}/rot_order"], g[f"{tag}/rots"]
    got = bvh.rotations_to_quat(rot_deg, order)
    assert isinstance(got, np.ndarray) and got.dtype == np.float32 and got.shape == want.shape
    assert_allclose(np.linalg.norm(got, axis=-1), 1.0, atol=1e-6)
    dots = np.sum(got * want, axis=-1)
    # degrees up to a few turns in fp32: the angle itself carries ~1e-5 rad of rounding before any arithmetic
    assert_allclose(np.abs(dots), 1.0, atol=2e-6)
    # the reference's own decisions: |dot| of consecutive reference frames below 1e-3 is a coin toss in fp32
    margin = np.abs(np.sum(want[1:] * want[:-1], axis=-1))
    decided = np.concatenate([np.ones((1,) + margin.shape[1:], bool), np.minimum.accumulate(margin > 1e-3, axis=0)])
    assert (dots[decided] > 0).all()
    assert decided.mean() > 0.9
    # sign-continuous along the frame axis, like the reference's output
    assert (np.sum(got[1:] * got[:-1], axis=-1) >= -1e-6).all()
    # the same chain as three separate ops of this package (from_euler -> unroll -> normalize) agrees bit for bit on
    # the cover and to rounding on the values
    sep = quat.normalize(quat.unroll(quat.from_euler(np.radians(rot_deg).astype(np.float32), np.tile(order, (rot_deg.shape[0], 1, 1))), 0))
    assert_allclose(got, sep, rtol=0, atol=2e-6)
    # drop-in for BVH.get_data given the reference's data dictionary
    data = {"rotations": rot_deg, "rot_order": order, "positions": np.zeros(rot_deg.shape, np.float32), "parents": np.arange(3),
            "offsets": np.zeros((3, 3)), "end_sites": np.zeros((0, 3)), "end_sites_parents": np.zeros(0, int)}
    rots, pos, parents, offsets, end_sites, end_sites_parents = bvh.get_data(data)
    assert_array_equal(rots, got)
    assert pos is data["positions"] and parents is data["parents"]
    t = bvh.rotations_to_quat(torch.from_numpy(rot_deg).cuda(), order)
    assert t.is_cuda and torch.equal(t.cpu(), torch.from_numpy(got))
