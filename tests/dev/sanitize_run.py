#!/usr/bin/env python
"""Small-shape pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tests/dev/sanitize_run.py

Checks results against the oracle too, so a sanitizer-clean run is also a correct one."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

from oracle import pymotion_oracle as orc  # noqa: E402
from pymotion_b200.ops import skeleton as sk  # noqa: E402
from pymotion_b200.rotations import dual_quat as dq  # noqa: E402
from pymotion_b200.rotations import quat  # noqa: E402
from pymotion_b200.topologies import parents_of, synth_numpy  # noqa: E402

KNOBS = [{}, {"PMB_FK_ROWS": "1"}, {"PMB_FK_ROWS": "1", "PMB_FK_STAGES": "3", "PMB_FK_BLOCKS_PER_SM": "1"},
         {"PMB_FK_MTRACKS": "0", "PMB_FK_ROWS": "0", "PMB_FK_TRACKS": "0"}, {"PMB_FK_GROUP": "8", "PMB_FK_WARPS": "4"},
         {"PMB_FK_TRACKS": "1"}, {"PMB_FK_TRACKS": "1", "PMB_FK_UL": "2", "PMB_FK_U": "1", "PMB_FK_NB": "4", "PMB_FK_WARPS_PER_SM": "2"},
         {"PMB_FK_TRACKS": "1", "PMB_FK_UL": "2", "PMB_FK_U": "2", "PMB_FK_NB": "4"},
         {"PMB_FK_MTRACKS": "1"}, {"PMB_FK_MTRACKS": "1", "PMB_FK_WARPS_PER_SM": "1"}, {"PMB_FK_MTRACKS": "0"}]


def set_knobs(knobs):
    from pymotion_b200 import _lib

    for k in [k for k in os.environ if k.startswith("PMB_")]:
        os.environ.pop(k)
    os.environ.update(knobs)
    if knobs:
        os.environ["PMB_EXPERIMENT"] = "1"
    _lib.load().pmb_reload_knobs()


def main():
    n_checked = 0
    rng = np.random.default_rng(3)
    shapes = [(parents_of(name), frames) for name, frames in (("body22", 1203), ("smplh52", 611), ("deep65", 395), ("chain3", 77))]
    shapes.append((np.array([0]), 70))  # a single joint
    for n_joints, frames in ((129, 45), (512, 40)):  # random joint orders, up to the largest skeleton the ABI takes
        par = np.zeros(n_joints, dtype=np.int64)
        for i in range(1, n_joints):
            par[i] = rng.integers(max(0, i - 6), i)
        shapes.append((par, frames))
    for par, frames in shapes:
        rot, gp, off = synth_numpy(frames, par, seed=frames)
        rot = (rot / np.linalg.norm(rot, axis=-1, keepdims=True)).astype(np.float32)
        want_pos, want_rotm = orc.fk(rot, gp, off, par)
        scale = max(1.0, float(np.abs(want_pos).max()))  # deep chains accumulate rounding
        for knobs in KNOBS:
            set_knobs(knobs)
            try:
                pos, rotm = sk.fk(rot, gp, off, par)
            except Exception as e:  # a FORCED variant may not fit the largest skeletons; the default policy must
                if knobs and ("fit" in str(e) or "variant" in str(e)):
                    continue
                raise
            np.testing.assert_allclose(pos, want_pos, rtol=1e-5, atol=1e-5 * scale)
            np.testing.assert_allclose(rotm, want_rotm, rtol=1e-5, atol=5e-5)
            n_checked += 1
        want_dq = orc.to_root_dual_quat(rot, gp, par, off)
        # the quaternion track kernel (forced, both tile-claim modes, with and without the one-step-ahead fetch) and the
        # thread-per-frame chain kernels it replaces by default
        for knobs in ({"PMB_DQ_TRACKS": "1", "PMB_FKQ_TRACKS": "1"},
                      {"PMB_DQ_TRACKS": "1", "PMB_FKQ_TRACKS": "1", "PMB_QT_DYNAMIC": "0", "PMB_QT_PIPE": "1", "PMB_QT_WARPS_PER_SM": "2"},
                      {"PMB_DQ_TRACKS": "1", "PMB_FKQ_TRACKS": "1", "PMB_QT_PIPE": "0", "PMB_QT_WARPS_PER_SM": "1"},
                      {"PMB_DQ_TRACKS": "1", "PMB_FKQ_TRACKS": "1", "PMB_QT_SHAPE": "1"},
                      {"PMB_DQ_TRACKS": "1", "PMB_FKQ_TRACKS": "1", "PMB_QT_SHAPE": "2", "PMB_QT_PIPE": "1"},
                      {"PMB_DQ_TRACKS": "0", "PMB_FKQ_TRACKS": "0"}, {}):
            set_knobs(knobs)
            try:
                p2, gq = sk.fk_quat(rot, gp, off, par)
                d = sk.to_root_dual_quat(rot, gp, par, off)
            except Exception as e:
                if knobs and ("fit" in str(e) or "variant" in str(e)):
                    continue
                raise
            np.testing.assert_allclose(p2, want_pos, rtol=1e-5, atol=1e-5 * scale)
            np.testing.assert_allclose(d, want_dq, rtol=1e-5, atol=5e-5 * scale)
            n_checked += 2
        t, r = sk.from_root_dual_quat(d, par)
        np.testing.assert_allclose(r, rot, rtol=1e-5, atol=5e-5)
        sk.from_global_rotations(gq, par)
        centred = (want_pos - want_pos[:, :1]).astype(np.float32)
        sk.from_root_positions(centred, par, off)
        for knobs in ({}, {"PMB_MIRROR_FUSED": "0"}, {"PMB_QT_WARPS_PER_SM": "1"}):  # fused mirror epilogue / two-kernel path
            set_knobs(knobs)
            for mode in ("all", "positions"):
                sk.mirror(rot, gp, par, off, mode=mode)
        set_knobs({})
        quat.unroll(rot, 0)
        dq.unroll(d, 0)
        dq.normalize(d)
        dq.is_unit(d)
        quat.to_matrix(rot)
        quat.slerp(rot, rot[::-1].copy(), 0.3)
        if len(par) <= 65:
            import pymotion_b200.io.bvh as bvh

            order = np.array([list("zyx")] * len(par))
            bvh.rotations_to_quat(np.degrees(rot[..., 1:]).astype(np.float32), order)
            # large host batches take the chunked pipeline; force it on this small one through fk_host
            hp, hr = sk.fk_host(rot, gp, off, par, chunk_frames=64)
            np.testing.assert_allclose(hp, want_pos, rtol=1e-5, atol=1e-5 * scale)
            hd = sk.to_root_dual_quat(rot, gp, par, off)   # NumPy in: the host pipeline of the dual-quaternion pair
            np.testing.assert_allclose(hd, want_dq, rtol=1e-5, atol=5e-5 * scale)
            sk.from_root_dual_quat(hd, par)
            hp, hq = sk.fk_quat_host(rot, gp, off, par, chunk_frames=64)
            np.testing.assert_allclose(hp, want_pos, rtol=1e-5, atol=1e-5 * scale)
            n_checked += 3
        n_checked += 12
    print(f"sanitize_run ok: {n_checked} kernel-family launches checked")


if __name__ == "__main__":
    main()
