"""Pins the oracle's restatement of from_root_positions / mirror against fixtures written by
the real reference (oracle/gen_golden.py::gen_ik).  CPU only; same arithmetic => bit-exact."""
import warnings

import numpy as np
import pytest
from numpy.testing import assert_array_equal

from oracle import pymotion_oracle as orc

SKELS = ("chain3", "body22", "smplh52", "deep65")


@pytest.fixture(autouse=True)
def _quiet():
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)  # sqrt of a rounding-negative 1 - dot inside from_to
        yield


@pytest.mark.parametrize("name", SKELS)
def test_from_root_positions(golden_ik, name):
    g = golden_ik
    got = orc.from_root_positions(g[f"{name}/centred"], g[f"{name}/parents"], g[f"{name}/offsets"])
    assert got.dtype == np.float64
    assert_array_equal(got, g[f"{name}/from_root_positions"])
    # leaves keep the identity (skeleton.py:129-132)
    par = g[f"{name}/parents"]
    leaves = [j for j in range(len(par)) if j not in set(par[1:].tolist())]
    assert_array_equal(got[:, leaves], np.tile([1.0, 0, 0, 0], (got.shape[0], len(leaves), 1)))


@pytest.mark.parametrize("name", SKELS)
def test_mirror_modes(golden_ik, name):
    g = golden_ik
    rot, gpos, par, off = (g[f"{name}/{k}"] for k in ("rot", "gpos", "parents", "offsets"))
    ends = g["end_sites"]
    keep = [a.copy() for a in (rot, gpos, off, ends)]
    for axis in ("XYZ" if name == "body22" else "Y"):
        out = orc.mirror(rot, gpos, par, off, ends, None, "all", axis)
        for key, val in zip(("rot", "gpos", "offsets", "ends"), out):
            assert_array_equal(val, g[f"{name}/mirror_all_{axis}/{key}"])
    out = orc.mirror(rot, gpos, par, off, None, None, "positions", "X")
    assert_array_equal(out[0], g[f"{name}/mirror_positions_X/rot"])
    assert_array_equal(out[1], g[f"{name}/mirror_positions_X/gpos"])
    assert out[2] is off and out[3] is None
    if name == "body22":
        for axis in "XZ":
            out = orc.mirror(rot, gpos, par, off, None, g["body22/joints_mapping"], "symmetry", axis)
            assert_array_equal(out[0], g[f"{name}/mirror_symmetry_{axis}/rot"])
            assert_array_equal(out[1], g[f"{name}/mirror_symmetry_{axis}/gpos"])
    for a, b in zip((rot, gpos, off, ends), keep):
        assert_array_equal(a, b)  # inputs untouched
    with pytest.raises(ValueError):
        orc.mirror(rot, gpos, par, off, mode="symmetry")
    with pytest.raises(ValueError):
        orc.mirror(rot, gpos, par, off, mode="nope")
    with pytest.raises(ValueError):
        orc.mirror(rot, gpos, par, off, axis="W")


def test_mirror_twice_is_identity(golden_ik):
    g = golden_ik
    rot, gpos, par, off = (g[f"body22/{k}"] for k in ("rot", "gpos", "parents", "offsets"))
    r1, g1, o1, _ = orc.mirror(rot, gpos, par, off, mode="all", axis="X")
    r2, g2, o2, _ = orc.mirror(r1, g1, par, o1, mode="all", axis="X")
    unit = rot / np.linalg.norm(rot, axis=-1, keepdims=True)
    np.testing.assert_allclose(np.abs(np.sum(r2 * unit, axis=-1)), 1.0, atol=1e-6)
    np.testing.assert_allclose(g2, gpos, atol=0)
    np.testing.assert_allclose(o2, off, atol=0)
