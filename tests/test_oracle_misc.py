"""Pins the oracle's restatement of ortho6d, center_of_mass, interpolate_positions and vector.normalize
(SURVEY 8f rank 3 tail / rank 4) against fixtures written by the real reference (gen_golden.py::gen_misc).
CPU only; same arithmetic => bit-exact."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from oracle import pymotion_oracle as orc

TAGS = ("f32", "f64")


def same(got, want):
    assert got.dtype == want.dtype and got.shape == want.shape
    assert_array_equal(got, want)


@pytest.mark.parametrize("tag", TAGS)
def test_ortho6d(golden_misc, tag):
    g = golden_misc
    same(orc.ortho6d_from_quat(g[f"{tag}/q"]), g[f"{tag}/o6_from_quat"])
    same(np.ascontiguousarray(orc.ortho6d_from_matrix(g[f"{tag}/m"])), g[f"{tag}/o6_from_matrix"])
    same(np.ascontiguousarray(orc.ortho6d_to_matrix(g[f"{tag}/o6"])), g[f"{tag}/o6_to_matrix"])
    same(orc.ortho6d_to_quat(g[f"{tag}/o6"]), g[f"{tag}/o6_to_quat"])
    # rotations/tests/test_ortho6d.py:14: matrix -> ortho6d -> matrix is the identity on rotation matrices
    back = orc.ortho6d_to_matrix(orc.ortho6d_from_matrix(g[f"{tag}/m"]))
    assert_allclose(back, g[f"{tag}/m"], atol=1e-6)


@pytest.mark.parametrize("tag", TAGS)
def test_center_of_mass(golden_misc, tag):
    g = golden_misc
    same(orc.center_of_mass(g[f"{tag}/joints"], g[f"{tag}/weights"]), g[f"{tag}/com"])
    same(orc.center_of_mass(g[f"{tag}/joints"], g[f"{tag}/weights_pf"]), g[f"{tag}/com_pf"])
    j = g[f"{tag}/joints"]
    same(orc.human_center_of_mass(j[..., 0:6, :], j[..., 6:10, :], j[..., 10:14, :], j[..., 14:18, :], j[..., 18:22, :]),
         g[f"{tag}/human_com"])
    # ops/tests/test_center_of_mass.py:15-40
    assert_allclose(orc.center_of_mass(np.array([[[1, 2, 3], [4, 5, 6], [7, 8, 9]]]), np.array([[0.3, 0.3, 0.4]])),
                    np.array([[4.3, 5.3, 6.3]]), atol=1e-6)
    same(orc.center_of_mass(g["hand/com_joints"], g["hand/com_weights"]), g["hand/com"])


@pytest.mark.parametrize("tag", TAGS)
def test_interpolate_positions(golden_misc, tag):
    g = golden_misc
    same(orc.interpolate_positions(g[f"{tag}/t_sample"], g[f"{tag}/t_orig"], g[f"{tag}/interp_p1"], 0), g[f"{tag}/interp_1"])
    same(orc.interpolate_positions(g[f"{tag}/t_sample"], g[f"{tag}/t_orig"], g[f"{tag}/interp_p2"], 1), g[f"{tag}/interp_2"])
    # samples on a knot reproduce the data
    assert_allclose(g[f"{tag}/interp_1"][-3:], g[f"{tag}/interp_p1"][[0, 7, 39]], atol=1e-6)


@pytest.mark.parametrize("tag", TAGS)
def test_vector_normalize(golden_misc, tag):
    g = golden_misc
    same(orc.vec_normalize(g[f"{tag}/vec"]), g[f"{tag}/vec_normalize"])
    same(orc.vec_normalize(g[f"{tag}/vec"], eps=1e-3), g[f"{tag}/vec_normalize_eps"])
    same(orc.vec_normalize(g[f"{tag}/vec5"]), g[f"{tag}/vec5_normalize"])


@pytest.mark.parametrize("tag", ["small", "body22", "long"])
def test_bvh_get_data_rotations(golden_bvh, tag):
    """The numeric chain of BVH.get_data (io/bvh.py:352-359) restated, against outputs of the REAL BVH class on an
    in-memory data dictionary (gen_golden.py::gen_bvh): bit-exact, unit length, sign-continuous along the frames."""
    g = golden_bvh
    got = orc.bvh_get_data_rotations(g[f"{tag}/rotations_deg"], g[f"{tag}/rot_order"])
    same(got, g[f"{tag}/rots"])
    assert_allclose(np.linalg.norm(got, axis=-1), 1.0, atol=1e-7)
    assert (np.sum(got[1:] * got[:-1], axis=-1) >= 0).all()
