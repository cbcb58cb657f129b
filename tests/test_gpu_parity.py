"""Parity of the CUDA path (through the C ABI) against the fixtures written by the
real reference and against the oracle on seeded inputs.  Needs a B200: -m gpu.

Tolerance: BASELINE.json's north_star asks for positions / rotmats within 1e-5
relative in fp32 of the reference's NumPy output; RTOL = ATOL = 1e-5 below.  The
reference tests' hand-written goldens are checked at their own atol (1e-6)."""
import warnings

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from oracle import oracle_c
from oracle import pymotion_oracle as orc
from pymotion_b200.topologies import TOPOLOGIES, parents_of, synth_numpy, synth_torch
from pymotion_b200 import _lib

pytestmark = pytest.mark.gpu

RTOL = ATOL = 1e-5
TOL = dict(rtol=RTOL, atol=ATOL)
SKELS = ("body22", "smplh52", "deep65")


@pytest.fixture(scope="module")
def sk():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import pymotion_b200.ops.skeleton as mod

    return mod


@pytest.fixture(scope="module")
def quat(sk):
    import pymotion_b200.rotations.quat as mod

    return mod


@pytest.fixture(scope="module")
def dquat(sk):
    import pymotion_b200.rotations.dual_quat as mod

    return mod


# ---------------------------------------------------------------- fk vs reference fixtures
@pytest.mark.parametrize("case", ["chain3_ident", "chain3_rot"])
def test_fk_chain3_goldens(sk, golden_fk, case):
    g = golden_fk
    rot = g[f"{case}/rot"].astype(np.float32)
    pos, rotm = sk.fk(rot, g["chain3/gpos"], g["chain3/offsets"], g["chain3/parents"])
    assert isinstance(pos, np.ndarray) and pos.dtype == np.float32 and pos.shape == (2, 3, 3)
    assert rotm.shape == (2, 3, 3, 3)
    assert_allclose(pos, g[f"{case}/pos"], **TOL)
    assert_allclose(rotm, g[f"{case}/rotm"], **TOL)
    assert_allclose(pos, g[f"{case}/hand_pos"], atol=1e-6)  # ops/tests/test_skeleton.py:270, :372
    if case == "chain3_rot":
        assert_allclose(rotm, g["chain3_rot/hand_rotm"], atol=1e-6)  # :373
    # per-frame offsets accepted (:267)
    pos2, rotm2 = sk.fk(rot, g["chain3/gpos"], np.tile(g["chain3/offsets"], (2, 1, 1)), g["chain3/parents"])
    assert_allclose(pos2, pos, **TOL)  # different kernels (row teams vs per-frame offsets): same answer, not the same bits
    assert_allclose(rotm2, rotm, **TOL)


def test_fk_nd_leading_dims(sk, golden_fk):
    g = golden_fk
    pos, rotm = sk.fk(g["chain3_nd/rot"].astype(np.float32), g["chain3_nd/gpos"], g["chain3/offsets"], g["chain3/parents"])
    assert pos.shape == (4, 3, 4, 3, 3) and rotm.shape == (4, 3, 4, 3, 3, 3)
    assert_allclose(pos, g["chain3_nd/pos"], **TOL)
    assert_allclose(rotm, g["chain3_nd/rotm"], **TOL)


@pytest.mark.parametrize("name", SKELS)
def test_fk_skeleton_fixtures(sk, golden_fk, name):
    g = golden_fk
    par, rot, gp, off = (g[f"{name}/{k}"] for k in ("parents", "rot", "gpos", "offsets"))
    pos, rotm = sk.fk(rot, gp, off, par)  # non-unit and one zero quaternion inside
    assert_allclose(pos, g[f"{name}/pos"], **TOL)
    assert_allclose(rotm, g[f"{name}/rotm"], **TOL)
    pos, rotm = sk.fk(rot, gp, g[f"{name}/offsets_pf"], par)  # per-frame offsets, offsets[:,0] != 0 ignored
    assert_allclose(pos, g[f"{name}/pos_pf"], **TOL)
    assert_allclose(rotm, g[f"{name}/rotm_pf"], **TOL)
    # torch CUDA tensors in -> torch CUDA tensors out, same values
    dev = torch.device("cuda")
    tp, tr = sk.fk(torch.from_numpy(rot).to(dev), torch.from_numpy(gp).to(dev), torch.from_numpy(off).to(dev),
                   torch.from_numpy(par))
    assert tp.is_cuda and tp.dtype == torch.float32 and tr.is_contiguous()
    assert_allclose(tp.cpu().numpy(), g[f"{name}/pos"], **TOL)
    # float64 in -> float64 out (computed in fp32, documented)
    p64, r64 = sk.fk(rot.astype(np.float64), gp.astype(np.float64), off.astype(np.float64), par)
    assert p64.dtype == np.float64
    assert_allclose(p64, g[f"{name}/pos_f64"], **TOL)
    assert_allclose(r64, g[f"{name}/rotm_f64"], **TOL)


def test_fk_edge_shapes(sk, golden_fk):
    g = golden_fk
    pos, rotm = sk.fk(g["unbatched/rot"], g["unbatched/gpos"], g["unbatched/offsets"], g["unbatched/parents"])
    assert pos.shape == (22, 3) and rotm.shape == (22, 3, 3)  # unbatched, parents[0] = -1
    assert_allclose(pos, g["unbatched/pos"], **TOL)
    assert_allclose(rotm, g["unbatched/rotm"], **TOL)
    pos, rotm = sk.fk(g["single/rot"], g["single/gpos"], g["single/offsets"], np.array([0]))
    assert_allclose(pos, g["single/pos"], **TOL)
    assert_allclose(rotm, g["single/rotm"], **TOL)
    pos, rotm = sk.fk(g["bcast/rot"], np.zeros((1, 3), dtype=np.float32), g["bcast/offsets"], g["body22/parents"])
    assert_allclose(pos, g["bcast/pos"], **TOL)
    assert_allclose(rotm, g["bcast/rotm"], **TOL)
    # empty batch
    pos, rotm = sk.fk(np.zeros((0, 22, 4), np.float32), np.zeros((0, 3), np.float32), g["bcast/offsets"], g["body22/parents"])
    assert pos.shape == (0, 22, 3) and rotm.shape == (0, 22, 3, 3)


def test_fk_input_not_mutated_and_errors(sk, golden_fk):
    g = golden_fk
    dev = torch.device("cuda")
    rot = torch.from_numpy(g["body22/rot"]).to(dev)
    keep = rot.clone()
    sk.fk(rot, torch.from_numpy(g["body22/gpos"]).to(dev), torch.from_numpy(g["body22/offsets"]).to(dev), g["body22/parents"])
    assert torch.equal(rot, keep)
    bad = g["body22/parents"].copy()
    bad[3] = 7
    with pytest.raises(ValueError):
        sk.fk(rot, torch.from_numpy(g["body22/gpos"]).to(dev), torch.from_numpy(g["body22/offsets"]).to(dev), bad)
    with pytest.raises(ValueError):
        sk.fk(rot[:, :5], torch.from_numpy(g["body22/gpos"]).to(dev), torch.from_numpy(g["body22/offsets"]).to(dev), g["body22/parents"])


# ---------------------------------------------------------------- fk vs oracle on seeded inputs, ragged sizes
@pytest.mark.parametrize("name", SKELS)
@pytest.mark.parametrize("n_frames", [1, 31, 32, 33, 127, 1000])
def test_fk_vs_numpy_oracle_ragged(sk, name, n_frames):
    par = parents_of(name)
    rot, gp, off = synth_numpy(n_frames, par, seed=100 + n_frames)
    pos, rotm = sk.fk(rot, gp, off, par)
    want_pos, want_rotm = orc.fk(rot, gp, off, par)
    assert_allclose(pos, want_pos, **TOL)
    assert_allclose(rotm, want_rotm, **TOL)


def test_fk_random_trees_vs_oracle(sk):
    rng = np.random.default_rng(5)
    for n_joints in (2, 7, 8, 40, 129, 512):
        par = np.zeros(n_joints, dtype=np.int64)
        for i in range(1, n_joints):
            par[i] = rng.integers(max(0, i - 6), i)  # bounded look-back keeps the slot count small
        rot, gp, off = synth_numpy(70, par, seed=n_joints)
        pos, rotm = sk.fk(rot, gp, off, par)
        want_pos, want_rotm = orc.fk(rot, gp, off, par)
        # deep chains accumulate: scale atol with depth * |offset|
        assert_allclose(pos, want_pos, rtol=RTOL, atol=ATOL * max(1.0, np.abs(want_pos).max()))
        assert_allclose(rotm, want_rotm, rtol=RTOL, atol=5 * ATOL)


def test_fk_full_size_1m_x_22_vs_c_oracle(sk):
    """BASELINE config 2 at full size against the C restatement (float32 locals,
    float64 chain), plus size-independent invariants on the whole batch."""
    par = parents_of("body22")
    dev = torch.device("cuda")
    n = 1_000_000
    rot, gp, off = synth_torch(n, par, dev)
    pos, rotm = sk.fk(rot, gp, off, par)
    torch.cuda.synchronize()
    # invariants on all 22M joints: orthonormal rotations, bone lengths preserved
    eye = torch.eye(3, device=dev)
    err = (rotm @ rotm.transpose(-1, -2) - eye).abs().amax().item()
    assert err < 1e-5
    assert (torch.linalg.det(rotm[::97]) - 1).abs().amax().item() < 1e-5
    ptorch = torch.as_tensor(par, device=dev)
    bone = (pos[:, 1:] - pos[:, ptorch[1:]]).norm(dim=-1)
    assert_allclose(bone.amax(0).cpu().numpy(), off[1:].norm(dim=-1).cpu().numpy(), rtol=1e-5, atol=1e-5)
    assert_allclose(bone.amin(0).cpu().numpy(), off[1:].norm(dim=-1).cpu().numpy(), rtol=1e-5, atol=1e-5)
    assert_allclose(pos[:, 0].cpu().numpy(), gp.cpu().numpy(), atol=0)
    # element-wise against the oracle, in four slabs to bound host memory
    off_h = off.cpu().numpy()
    for lo in range(0, n, 250_000):
        sl = slice(lo, lo + 250_000)
        want_pos, want_rotm = oracle_c.fk(rot[sl].cpu().numpy(), gp[sl].cpu().numpy(), off_h, par)
        assert_allclose(pos[sl].cpu().numpy(), want_pos, **TOL)
        assert_allclose(rotm[sl].cpu().numpy(), want_rotm, **TOL)


@pytest.mark.parametrize("name", ["deep65", "smplh52"])
def test_fk_4m_invariants_and_slabs(sk, name):
    """BASELINE config 4 (4M x 65, deep hierarchy) and the per-GPU shard of config 5 (4M x 52): invariants on the
    full batch, oracle on the first / last / a middle 64k-frame window."""
    par = parents_of(name)
    dev = torch.device("cuda")
    n = 4_000_000
    rot, gp, off = synth_torch(n, par, dev, seed=4321)
    pos, rotm = sk.fk(rot, gp, off, par)
    torch.cuda.synchronize()
    ptorch = torch.as_tensor(par, device=dev)
    want_len = off[1:].norm(dim=-1)
    worst = 0.0
    for lo in range(0, n, 500_000):  # slabs keep the temporaries small
        sl = slice(lo, lo + 500_000)
        bone = (pos[sl, 1:] - pos[sl][:, ptorch[1:]]).norm(dim=-1)
        worst = max(worst, (bone - want_len).abs().amax().item())
        r = rotm[sl]
        worst_r = (r @ r.transpose(-1, -2) - torch.eye(3, device=dev)).abs().amax().item()
        assert worst_r < 2e-5
    assert worst < 2e-5
    off_h = off.cpu().numpy()
    for lo in (0, 1_777_777, n - 65_536):
        sl = slice(lo, lo + 65_536)
        want_pos, want_rotm = oracle_c.fk(rot[sl].cpu().numpy(), gp[sl].cpu().numpy(), off_h, par)
        assert_allclose(pos[sl].cpu().numpy(), want_pos, **TOL)
        assert_allclose(rotm[sl].cpu().numpy(), want_rotm, **TOL)


def assert_sign_convention(grot, want_q, rotm):
    """fk_quat returns the quaternion quat.from_matrix(fk rotmats) returns, SIGN INCLUDED (the cover feeds unroll /
    interpolation downstream).  The only entries allowed to differ are those where a branch test of from_matrix
    (quat.py:111-155: m22 < 0, m00 > m11, m00 < -m11) is within float32 rounding of a tie -- measured on B200: 4 of
    22 M entries at 1M x 22, 4 of 20.8 M at 400k x 52, 0 of 19.5 M at 300k x 65, every one with a tie margin below 1.4e-6
    (profiles/r2_fkq_sign_stats.jsonl)."""
    m = np.asarray(rotm, dtype=np.float64)
    flipped = np.sum(np.asarray(grot, dtype=np.float64) * want_q, axis=-1) < 0
    m22, m00, m11 = m[..., 2, 2], m[..., 0, 0], m[..., 1, 1]
    margin = np.minimum(np.abs(m22), np.where(m22 < 0, np.abs(m00 - m11), np.abs(m00 + m11)))
    assert flipped.mean() <= 1e-4, flipped.mean()
    assert (margin[flipped] < 1e-5).all(), margin[flipped].max()


def test_fk_quat_matches_from_matrix_of_fk(sk, quat, golden_fk):
    g = golden_fk
    for name in SKELS:
        par, rot, gp, off = (g[f"{name}/{k}"] for k in ("parents", "rot", "gpos", "offsets"))
        pos, grot = sk.fk_quat(rot, gp, off, par)
        assert_allclose(pos, g[f"{name}/pos"], **TOL)
        want = orc.quat_from_matrix(g[f"{name}/rotm"])
        # same rotation; sign may flip only where the branch test of from_matrix is within rounding of a tie
        dots = np.abs(np.sum(grot * want, axis=-1))
        assert_allclose(dots, 1.0, atol=1e-5)
        assert_sign_convention(grot, want, g[f"{name}/rotm"])


def test_fk_on_side_stream(sk, golden_fk):
    g = golden_fk
    dev = torch.device("cuda")
    par = g["body22/parents"]
    rot, gp, off = (torch.from_numpy(g[f"body22/{k}"]).to(dev) for k in ("rot", "gpos", "offsets"))
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        pos, rotm = sk.fk(rot, gp, off, par)
    s.synchronize()
    assert_allclose(pos.cpu().numpy(), g["body22/pos"], **TOL)


def test_concurrent_threads_on_their_own_streams(sk):
    """calls from several host threads, each on its own CUDA stream, are independent (ctypes drops the
    GIL for the duration of a call, so the host side of the library really runs concurrently): every thread gets the
    bits the same call gives alone, and a failing call in one thread does not disturb the others' error state."""
    import threading

    dev = torch.device("cuda")
    jobs = []
    for k, (name, n_frames) in enumerate((("body22", 3001), ("smplh52", 1500), ("deep65", 777), ("body22", 64))):
        par = parents_of(name)
        rot, gp, off = synth_torch(n_frames, par, dev, seed=900 + k)
        pos, rotm = sk.fk(rot, gp, off, par)
        dq = sk.to_root_dual_quat(rot, gp, par, off)
        jobs.append((par, rot, gp, off, pos.clone(), rotm.clone(), dq.clone()))
    torch.cuda.synchronize()
    failures = []

    def work(k):
        par, rot, gp, off, want_pos, want_rotm, want_dq = jobs[k]
        stream = torch.cuda.Stream()
        try:
            with torch.cuda.stream(stream):
                for it in range(25):
                    pos, rotm = sk.fk(rot, gp, off, par)
                    dq = sk.to_root_dual_quat(rot, gp, par, off)
                    if k == 3 and it % 5 == 0:  # an error path in the middle of everyone else's launches
                        bad = par.copy()
                        bad[3] = 7
                        with pytest.raises(ValueError):
                            sk.fk(rot, gp, off, bad)
                    stream.synchronize()
                    if not (torch.equal(pos, want_pos) and torch.equal(rotm, want_rotm) and torch.equal(dq, want_dq)):
                        failures.append((k, it))
        except Exception as e:  # noqa: BLE001 -- reported by the main thread
            failures.append((k, repr(e)))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not failures, failures


# ---------------------------------------------------------------- dual quaternions
@pytest.mark.parametrize("case", ["chain3_ident", "chain3_rot"])
def test_dq_chain3_goldens(sk, dquat, quat, golden_dq, case):
    g = golden_dq
    par, off, gp = g["chain3/parents"], g["chain3/offsets"], g["chain3/gpos"]
    rot = g[f"{case}/rot"].astype(np.float32)
    dq = sk.to_root_dual_quat(rot, gp, par, off)
    assert dq.shape == (2, 3, 8) and dq.dtype == np.float32
    assert_allclose(dq, g[f"{case}/dq"], **TOL)
    rr, tt = dquat.to_rotation_translation(dq)
    assert_allclose(tt[:, 1:], g[f"{case}/hand_root_trans"][:, 1:], atol=1e-6)  # test_skeleton.py:59, :170
    assert_allclose(tt[:, 0], gp, atol=1e-6)
    if case == "chain3_rot":
        assert_allclose(rr, orc.quat_from_matrix(g["chain3_rot/hand_root_rotm"]), atol=1e-6)  # :169
    trans, rots = sk.from_root_dual_quat(dq, par)  # (translations, rotations) order, skeleton.py:204
    assert trans.shape == (2, 3, 3) and rots.shape == (2, 3, 4)
    assert_allclose(rots, rot, atol=1e-6)  # :71, :184
    assert_allclose(trans[:, 1:], np.tile(off[1:], (2, 1, 1)), atol=1e-6)
    assert_allclose(trans[:, 0], gp, atol=1e-6)


@pytest.mark.parametrize("name", SKELS)
def test_dq_skeleton_fixtures(sk, golden_dq, name):
    g = golden_dq
    par, rot, gp, off = (g[f"{name}/{k}"] for k in ("parents", "rot", "gpos", "offsets"))
    dq = sk.to_root_dual_quat(rot, gp, par, off)  # slightly non-unit quaternions: no normalisation inside
    assert_allclose(dq, g[f"{name}/dq"], **TOL)
    trans, rots = sk.from_root_dual_quat(g[f"{name}/dq"].astype(np.float32), par)
    assert_allclose(trans, g[f"{name}/back_trans_f32in"], **TOL)
    assert_allclose(rots, g[f"{name}/back_rot_f32in"], **TOL)
    # N-D leading dims use shape[-2] as the joint count (documented deviation from skeleton.py:228)
    dq_nd = sk.to_root_dual_quat(rot.reshape((1,) + rot.shape), gp.reshape((1,) + gp.shape), par, off)
    assert_allclose(dq_nd[0], dq, atol=0)


@pytest.mark.parametrize("n_frames", [1, 33, 64, 65, 1000])
def test_dq_vs_oracle_ragged(sk, n_frames):
    for name in SKELS:
        par = parents_of(name)
        rot, gp, off = synth_numpy(n_frames, par, seed=300 + n_frames)
        dq = sk.to_root_dual_quat(rot, gp, par, off)
        want = orc.to_root_dual_quat(rot, gp, par, off)
        assert_allclose(dq, want, **TOL)
        trans, rots = sk.from_root_dual_quat(want.astype(np.float32), par)
        wt, wr = orc.from_root_dual_quat(want, par)
        assert_allclose(trans, wt, **TOL)
        assert_allclose(rots, wr, **TOL)


def test_dq_asserts_like_reference(sk, golden_dq):
    g = golden_dq
    off = g["body22/offsets"].copy()
    off[0, 2] = 1.0
    with pytest.raises(AssertionError):  # ops/skeleton.py:227
        sk.to_root_dual_quat(g["body22/rot"], g["body22/gpos"], g["body22/parents"], off)
    with pytest.raises(AssertionError):  # per-frame offsets are rejected by the same assert
        sk.to_root_dual_quat(g["body22/rot"], g["body22/gpos"], g["body22/parents"], np.tile(g["body22/offsets"], (37, 1, 1)))


def test_dq_round_trip_1m_x_22(sk):
    """BASELINE config 3 at full size: to_root_dual_quat -> from_root_dual_quat gives
    back (offsets | global_pos, rotations) -- the identity the reference test checks
    (test_skeleton.py:70-77) -- plus the oracle on slabs."""
    par = parents_of("body22")
    dev = torch.device("cuda")
    n = 1_000_000
    rot, gp, off = synth_torch(n, par, dev, seed=99)
    dq = sk.to_root_dual_quat(rot, gp, par, off)
    trans, rots = sk.from_root_dual_quat(dq, par)
    torch.cuda.synchronize()
    assert (rots - rot).abs().amax().item() < 1e-5
    assert (trans[:, 1:] - off[1:]).abs().amax().item() < 1e-5
    assert (trans[:, 0] - gp).abs().amax().item() < 1e-5
    # unit dual quaternions: |q_r| = 1, q_r . q_d = 0
    assert (dq[..., :4].norm(dim=-1) - 1).abs().amax().item() < 1e-5
    assert (dq[..., :4] * dq[..., 4:]).sum(-1).abs().amax().item() < 1e-5
    off_h = off.cpu().numpy()
    for lo in (0, 500_000, n - 100_000):
        sl = slice(lo, lo + 100_000)
        want = oracle_c.to_root_dual_quat(rot[sl].cpu().numpy(), gp[sl].cpu().numpy(), par, off_h)
        assert_allclose(dq[sl].cpu().numpy(), want, **TOL)
        wt, wr = oracle_c.from_root_dual_quat(want, par)
        assert_allclose(trans[sl].cpu().numpy(), wt, **TOL)
        assert_allclose(rots[sl].cpu().numpy(), wr, **TOL)


@pytest.mark.parametrize("name", ["deep65", "smplh52"])
def test_quaternion_walks_4m(sk, name):
    """The quaternion track kernel at the full sizes of BASELINE configs 4 / 5 (4M x 65, 4M x 52 per GPU), through
    size-independent properties on EVERY frame -- dual quaternions are unit and decode back to the inputs (the identity the
    reference test checks, test_skeleton.py:70-77); fk_quat gives the positions of fk and quaternions whose matrices are fk's
    matrices; mirroring twice gives the rotations back -- plus the C oracle on the first / a middle / the last 64k-frame window."""
    par = parents_of(name)
    dev = torch.device("cuda")
    n = 4_000_000
    rot, gp, off = synth_torch(n, par, dev, seed=777)
    dq = sk.to_root_dual_quat(rot, gp, par, off)
    assert "qtracks_kernel" in _lib.load().pmb_last_variant().decode()
    trans, rots = sk.from_root_dual_quat(dq, par)
    ptorch = torch.as_tensor(par, device=dev)
    root_child = ptorch == 0
    for lo in range(0, n, 500_000):  # slabs keep the temporaries small
        sl = slice(lo, lo + 500_000)
        assert (rots[sl] - rot[sl]).abs().amax().item() < 1e-5
        assert (trans[sl, 1:] - off[1:]).abs().amax().item() < 2e-5
        assert (trans[sl, 0] - gp[sl]).abs().amax().item() < 1e-5
        assert (dq[sl][..., :4].norm(dim=-1) - 1).abs().amax().item() < 1e-5
        assert (dq[sl][..., :4] * dq[sl][..., 4:]).sum(-1).abs().amax().item() < 2e-5
    off_h = off.cpu().numpy()
    for lo in (0, 1_777_777, n - 65_536):
        sl = slice(lo, lo + 65_536)
        want = oracle_c.to_root_dual_quat(rot[sl].cpu().numpy(), gp[sl].cpu().numpy(), par, off_h)
        assert_allclose(dq[sl].cpu().numpy(), want, **TOL)
    del dq, trans, rots, root_child
    torch.cuda.empty_cache()

    pos, rotm = sk.fk(rot, gp, off, par)
    qpos, grot = sk.fk_quat(rot, gp, off, par)
    assert "qtracks_kernel" in _lib.load().pmb_last_variant().decode()
    from pymotion_b200.rotations import quat as gquat

    for lo in range(0, n, 500_000):
        sl = slice(lo, lo + 500_000)
        assert (qpos[sl] - pos[sl]).abs().amax().item() < 2e-5
        assert (gquat.to_matrix(grot[sl]) - rotm[sl]).abs().amax().item() < 2e-5
    del pos, rotm, qpos, grot
    torch.cuda.empty_cache()

    zero = torch.zeros(3, device=dev)
    r1, _, o1, _ = sk.mirror(rot, zero, par, off, mode="all", axis="X")
    assert "MODE=3" in _lib.load().pmb_last_variant().decode()
    r2, _, o2, _ = sk.mirror(r1, zero, par, o1, mode="all", axis="X")
    assert torch.equal(o2, off)
    for lo in range(0, n, 500_000):
        sl = slice(lo, lo + 500_000)
        assert ((r2[sl] * rot[sl]).sum(-1).abs() - 1).abs().amax().item() < 1e-5


def test_from_global_rotations(sk, golden_dq):
    g = golden_dq
    got = sk.from_global_rotations(g["fgr/global"], g["fgr/parents"])
    assert_allclose(got, g["fgr/local"], **TOL)
    # fk_quat -> from_global_rotations recovers the (normalised) local rotations up to sign
    par = parents_of("body22")
    rot, gp, off = synth_numpy(50, par, seed=77)
    _, grot = sk.fk_quat(rot, gp, off, par)
    local = sk.from_global_rotations(grot, par)
    assert_allclose(np.abs(np.sum(local * rot, axis=-1)), 1.0, atol=1e-5)


def test_random_trees_every_chain_op(sk):
    """Random joint orders (2 .. 512 joints, branches up to 6 joints back) through the other walks of the path:
    fk_quat, to / from_root_dual_quat, from_global_rotations.  Deep chains accumulate rounding, so the tolerance
    scales with the size of the values."""
    rng = np.random.default_rng(11)
    for n_joints in (2, 7, 8, 40, 129, 512):
        par = np.zeros(n_joints, dtype=np.int64)
        for i in range(1, n_joints):
            par[i] = rng.integers(max(0, i - 6), i)
        rot, gp, off = synth_numpy(70, par, seed=1000 + n_joints)
        rot = (rot / np.linalg.norm(rot, axis=-1, keepdims=True)).astype(np.float32)
        want_pos, want_rotm = orc.fk(rot, gp, off, par)
        pos, grot = sk.fk_quat(rot, gp, off, par)
        assert_allclose(pos, want_pos, rtol=RTOL, atol=ATOL * max(1.0, np.abs(want_pos).max()))
        assert_allclose(orc.quat_to_matrix(grot.astype(np.float64)), want_rotm, rtol=RTOL, atol=5 * ATOL)
        local = sk.from_global_rotations(grot, par)
        assert_allclose(np.abs(np.sum(local * rot, axis=-1)), 1.0, atol=5 * ATOL)
        dq = sk.to_root_dual_quat(rot, gp, par, off)
        want = orc.to_root_dual_quat(rot, gp, par, off)
        assert_allclose(dq, want, rtol=RTOL, atol=5 * ATOL * max(1.0, np.abs(want).max()))
        trans, rots = sk.from_root_dual_quat(want.astype(np.float32), par)
        wt, wr = orc.from_root_dual_quat(want.astype(np.float32), par)
        assert_allclose(trans, wt, rtol=RTOL, atol=5 * ATOL * max(1.0, np.abs(want).max()))
        assert_allclose(rots, wr, rtol=RTOL, atol=5 * ATOL)


@pytest.mark.parametrize("n_frames", [1, 33, 1000, 5001])
def test_single_joint_skeleton_every_op(sk, n_frames):
    """A skeleton of ONE joint (parents = [0]) through every op of the path: the element-per-thread kernels divide by
    the joint count with a multiply-high, and 1 is the divisor that has no 32-bit magic."""
    par = np.array([0])
    rng = np.random.default_rng(n_frames)
    rot = rng.normal(size=(n_frames, 1, 4)).astype(np.float32)
    rot /= np.linalg.norm(rot, axis=-1, keepdims=True)
    gp = rng.normal(size=(n_frames, 3)).astype(np.float32)
    off = np.zeros((1, 3), np.float32)
    pos, rotm = sk.fk(rot, gp, off, par)
    wp, wr = orc.fk(rot, gp, off, par)
    assert_allclose(pos, wp, **TOL)
    assert_allclose(rotm, wr, **TOL)
    pos_q, grot = sk.fk_quat(rot, gp, off, par)
    assert_allclose(pos_q, wp, **TOL)
    assert_allclose(np.abs(np.sum(grot * rot, axis=-1)), 1.0, atol=1e-5)
    dq = sk.to_root_dual_quat(rot, gp, par, off)
    want = orc.to_root_dual_quat(rot, gp, par, off)
    assert_allclose(dq, want, **TOL)
    trans, rots = sk.from_root_dual_quat(want.astype(np.float32), par)
    wt, wq = orc.from_root_dual_quat(want, par)
    assert_allclose(trans, wt, **TOL)
    assert_allclose(rots, wq, **TOL)
    assert_allclose(sk.from_global_rotations(rot, par), orc.from_global_rotations(rot, par), **TOL)
    r, t, o, _ = sk.mirror(rot, gp, par, off, mode="all", axis="X")
    mr, mt, mo, _ = orc.mirror(rot, gp, par, off, mode="all", axis="X")
    assert_allclose(np.abs(np.sum(r * mr, axis=-1)), 1.0, atol=1e-5)
    assert_allclose(t, mt, atol=0)
    assert_allclose(o, mo, atol=0)


# ---------------------------------------------------------------- primitives
def test_quat_primitives_vs_reference(quat, dquat, golden_quat):
    g = golden_quat
    q0, q1, v, qu, t = (g[f"f32/{k}"] for k in ("q0", "q1", "v", "qu", "t"))
    assert_allclose(quat.mul(q0, q1), g["f32/mul"], **TOL)
    assert_allclose(quat.mul(q0[:, :1], q1), g["f32/mul_bcast"], **TOL)
    assert_allclose(quat.mul_vec(q0, v), g["f32/mul_vec"], rtol=1e-5, atol=5e-5)  # |q|^2 up to ~20 scales the result
    assert_allclose(quat.length(q0), g["f32/length"], **TOL)
    assert quat.length(q0).shape == (3, 41)
    assert_allclose(quat.normalize(q0), g["f32/normalize"], **TOL)
    assert_allclose(quat.normalize(q0, eps=1e-2), g["f32/normalize_eps"], **TOL)
    assert_allclose(quat.conjugate(q0), g["f32/conjugate"], atol=0)
    assert_allclose(quat.inverse(q0), g["f32/inverse"], atol=0)
    assert_allclose(quat.to_matrix(q0), g["f32/to_matrix"], rtol=1e-5, atol=5e-5)
    assert quat.to_matrix(q0).shape == (3, 41, 3, 3)
    assert_allclose(quat.to_matrix(qu), g["f32/unit_matrix"], **TOL)
    assert_allclose(quat.from_matrix(g["f32/unit_matrix"].astype(np.float32)), g["f32/from_matrix"], **TOL)
    assert_allclose(quat.from_matrix(g["branches/matrix"].astype(np.float32)), g["branches/quat"], **TOL)
    dq = dquat.from_rotation_translation(qu, t)
    assert_allclose(dq, g["f32/dq"], **TOL)
    rr, tt = dquat.to_rotation_translation(g["f32/dq"].astype(np.float32))
    assert_allclose(rr, g["f32/dq_rot"], **TOL)
    assert_allclose(tt, g["f32/dq_trans"], **TOL)
    assert_allclose(tt, t, atol=1e-5)  # test_dual_quat.py:37-40 round trip
    assert_allclose(dquat.from_translation(t), g["f32/dq_from_translation"], **TOL)
    # hand-written goldens of rotations/tests/test_quat.py
    assert_allclose(quat.mul(g["hand/qa"], g["hand/qb"]), g["hand/mul_ab"], atol=1e-6)
    assert_allclose(quat.mul(g["hand/qb"], g["hand/qa"]), g["hand/mul_ba"], atol=1e-6)
    assert_allclose(quat.mul_vec(g["hand/qa"], g["hand/v"]), g["hand/mul_vec"], atol=1e-6)
    assert_allclose(quat.to_matrix(g["hand/qa"]), g["hand/matrix"], atol=1e-6)
    assert_allclose(quat.from_matrix(g["hand/matrix"]), g["hand/qa"], atol=1e-6)


def test_quat_large_flat_batch(quat):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((300_001, 4)).astype(np.float32)
    b = rng.standard_normal((300_001, 4)).astype(np.float32)
    assert_allclose(quat.mul(a, b), orc.quat_mul(a, b), rtol=1e-5, atol=1e-5)
    assert_allclose(quat.normalize(a), orc.quat_normalize(a), **TOL)
    m = orc.quat_to_matrix(orc.quat_normalize(a)).astype(np.float32)
    assert_allclose(quat.from_matrix(m), orc.quat_from_matrix(m), **TOL)


def test_host_buffer_entry_point(sk, golden_fk):
    """pmb_fk_f32_host: pinned host buffers in, host buffers out, chunked pipeline."""
    from pymotion_b200 import _lib

    lib = _lib.load()
    par = parents_of("body22")
    rot, gp, off = synth_numpy(10_000, par, seed=8)
    rot_t, gp_t = torch.from_numpy(rot).pin_memory(), torch.from_numpy(gp).pin_memory()
    pos = torch.empty((10_000, 22, 3), dtype=torch.float32).pin_memory()
    rotm = torch.empty((10_000, 22, 3, 3), dtype=torch.float32).pin_memory()
    rc = lib.pmb_fk_f32_host(rot_t.data_ptr(), gp_t.data_ptr(), off.ctypes.data, par.ctypes.data, 10_000, 22,
                             pos.data_ptr(), rotm.data_ptr(), 3000)  # ragged chunks: 3008, 3008, 3008, 976
    _lib.check(rc)
    want_pos, want_rotm = orc.fk(rot, gp, off, par)
    assert_allclose(pos.numpy(), want_pos, **TOL)
    assert_allclose(rotm.numpy(), want_rotm, **TOL)
    lib.pmb_release_workspace()


# ---------------------------------------------------------------- every kernel variant the host can select
@pytest.mark.parametrize("knobs", [
    {"PMB_FK_GROUP": "0", "PMB_FK_WARPS": "4"},   # whole-row stage + TMA bulk store (default for small skeletons)
    {"PMB_FK_GROUP": "0", "PMB_FK_WARPS": "5"},
    {"PMB_FK_GROUP": "0", "PMB_FK_WARPS": "2"},
    {"PMB_FK_GROUP": "32", "PMB_FK_WARPS": "4"},  # padded stage, periodic copy-out
    {"PMB_FK_GROUP": "16", "PMB_FK_WARPS": "4"},
    {"PMB_FK_GROUP": "8", "PMB_FK_WARPS": "4"},
    {"PMB_FK_GROUP": "8", "PMB_FK_WARPS": "1"},
    {"PMB_FK_GROUP": "8", "PMB_FK_WARPS": "4", "PMB_FK_BLOCKS_PER_SM": "1"},
])
@pytest.mark.parametrize("name,n_frames", [("body22", 4099), ("smplh52", 2050), ("deep65", 1031), ("chain3", 777)])
def test_fk_every_variant(sk, set_knobs, knobs, name, n_frames):
    """The launch heuristic picks one variant per shape; force each of them (ragged frame counts so the
    remainder tile and remainder group paths run) and compare with the oracle.  A variant that does not
    fit the shape (e.g. whole rows of a 65-joint skeleton with 5 warps) must refuse, not fall back."""
    set_knobs(knobs)
    par = parents_of(name)
    rot, gp, off = synth_numpy(n_frames, par, seed=len(par) + n_frames)
    want_pos, want_rotm = orc.fk(rot, gp, off, par)
    try:
        pos, rotm = sk.fk(rot, gp, off, par)
    except ValueError as e:
        assert "select no available variant" in str(e)
        return
    assert_allclose(pos, want_pos, **TOL)
    assert_allclose(rotm, want_rotm, **TOL)
    # per-frame offsets and quaternion output share the kernel: the two extremes of the group range
    if knobs["PMB_FK_GROUP"] in ("0", "8") and knobs["PMB_FK_WARPS"] == "4":
        off_pf = np.tile(off, (n_frames, 1, 1)) * np.linspace(0.5, 1.5, n_frames, dtype=np.float32)[:, None, None]
        want_pos, want_rotm = orc.fk(rot, gp, off_pf, par)
        try:
            pos, rotm = sk.fk(rot, gp, off_pf, par)
            assert_allclose(pos, want_pos, **TOL)
            assert_allclose(rotm, want_rotm, **TOL)
            pos, grot = sk.fk_quat(rot, gp, off, par)
            want_q = orc.quat_from_matrix(orc.fk(rot, gp, off, par)[1])
            assert_allclose(np.abs(np.sum(grot * want_q, axis=-1)), 1.0, atol=1e-5)
        except ValueError as e:
            assert "select no available variant" in str(e)


@pytest.mark.parametrize("knobs", [
    {"PMB_FK_ROWS": "1"},                                             # ring depth picked by the host
    {"PMB_FK_ROWS": "1", "PMB_FK_STAGES": "2"},
    {"PMB_FK_ROWS": "1", "PMB_FK_STAGES": "3"},
    {"PMB_FK_ROWS": "1", "PMB_FK_STAGES": "4"},
    {"PMB_FK_ROWS": "1", "PMB_FK_STAGES": "2", "PMB_FK_BLOCKS_PER_SM": "1"},   # many tiles per team
    {"PMB_FK_ROWS": "0"},                                             # the thread-per-frame chain kernel
])
@pytest.mark.parametrize("name,n_frames", [("body22", 40_003), ("smplh52", 20_050), ("deep65", 10_031), ("chain3", 777),
                                           ("body22", 31)])
def test_fk_row_team_kernel(sk, set_knobs, knobs, name, n_frames):
    """The row-team kernel (three warps per tile, loader and drainer threads): forced for every ring depth,
    with ragged frame counts (remainder tile through the drainer's 16-byte tail path) and with one block per
    SM so that every team walks several tiles (stage reuse, ring wrap-around)."""
    set_knobs(knobs)
    par = parents_of(name)
    rot, gp, off = synth_numpy(n_frames, par, seed=3 * len(par) + n_frames)
    want_pos, want_rotm = orc.fk(rot, gp, off, par)
    for _ in range(2):  # twice: the second launch overwrites the same outputs
        pos, rotm = sk.fk(rot, gp, off, par)
        assert_allclose(pos, want_pos, **TOL)
        assert_allclose(rotm, want_rotm, **TOL)


@pytest.mark.parametrize("knobs", [{}, {"PMB_FKQ_TRACKS": "0"}, {"PMB_FKQ_GROUP": "8"}, {"PMB_FKQ_GROUP": "16"}, {"PMB_FKQ_GROUP": "24"},
                                   {"PMB_FKQ_BLOCKS_PER_SM": "1"}, {"PMB_FKQ_MATRIX": "1"},
                                   {"PMB_FKQ_TRACKS": "1"}, {"PMB_FKQ_TRACKS": "1", "PMB_QT_WARPS_PER_SM": "1"},
                                   {"PMB_FKQ_TRACKS": "1", "PMB_QT_PIPE": "0", "PMB_QT_DYNAMIC": "0"},
                                   {"PMB_FKQ_TRACKS": "1", "PMB_QT_SHAPE": "1"}, {"PMB_FKQ_TRACKS": "1", "PMB_QT_SHAPE": "2"},
                                   {"PMB_FKQ_TRACKS": "1", "PMB_QT_SHAPE": "2", "PMB_QT_PIPE": "1", "PMB_QT_WARPS_PER_SM": "2"},
                                   {"PMB_FKQ_TRACKS": "1", "PMB_QT_PIPE": "1", "PMB_QT_WARPS_PER_SM": "3"}])
@pytest.mark.parametrize("name,n_frames", [("body22", 40_003), ("smplh52", 6_050), ("deep65", 5_031), ("chain3", 777),
                                           ("body40", 2_049)])
def test_fk_quat_every_variant(sk, set_knobs, knobs, name, n_frames):
    """fk_quat: the quaternion-chain kernel for every flush group (ragged frame counts: remainder tile and
    remainder group), one block per SM (several tiles per warp), and the older matrix-path kernel.  Positions
    at the hot-path tolerance; rotations equal to quat.from_matrix(fk rotmats) with the reference's sign
    convention except where from_matrix's branch test is within rounding of a tie."""
    set_knobs(knobs)
    par = parents_of(name)
    rot, gp, off = synth_numpy(n_frames, par, seed=11 * len(par) + n_frames)
    want_pos, want_rotm = orc.fk(rot, gp, off, par)
    want_q = orc.quat_from_matrix(want_rotm)
    pos, grot = sk.fk_quat(rot, gp, off, par)
    assert_allclose(pos, want_pos, **TOL)
    dots = np.sum(grot * want_q, axis=-1)
    assert_allclose(np.abs(dots), 1.0, atol=1e-5)
    assert_sign_convention(grot, want_q, want_rotm)
    assert_allclose(np.linalg.norm(grot, axis=-1), 1.0, atol=1e-5)


@pytest.mark.parametrize("knobs", [
    {"PMB_FK_TRACKS": "1"},                                                   # two tracks, three boxes, tiles of 10 frames
    {"PMB_FK_TRACKS": "1", "PMB_FK_U": "1", "PMB_FK_NB": "2"},
    {"PMB_FK_TRACKS": "1", "PMB_FK_U": "1", "PMB_FK_NB": "3", "PMB_FK_FR": "8"},
    {"PMB_FK_TRACKS": "1", "PMB_FK_U": "2", "PMB_FK_NB": "2"},
    {"PMB_FK_TRACKS": "1", "PMB_FK_U": "2", "PMB_FK_NB": "4", "PMB_FK_FR": "8", "PMB_FK_WARPS_PER_SM": "1"},  # many tiles per warp
    {"PMB_FK_TRACKS": "1", "PMB_FK_U": "2", "PMB_FK_NB": "3", "PMB_FK_WARPS_PER_SM": "3"},
    {"PMB_FK_TRACKS": "1", "PMB_FK_UL": "2", "PMB_FK_U": "1", "PMB_FK_NB": "3"},   # two lane groups, tiles of 5 frames
    {"PMB_FK_TRACKS": "1", "PMB_FK_UL": "2", "PMB_FK_U": "1", "PMB_FK_NB": "2", "PMB_FK_WARPS_PER_SM": "1"},
    {"PMB_FK_TRACKS": "1", "PMB_FK_UL": "2", "PMB_FK_U": "2", "PMB_FK_NB": "4"},   # four tracks
    {"PMB_FK_TRACKS": "1", "PMB_FK_UL": "2", "PMB_FK_U": "1", "PMB_FK_NB": "4", "PMB_FK_FR": "4"},
])
@pytest.mark.parametrize("name,n_frames", [("body22", 40_003), ("smplh52", 20_051), ("deep65", 10_031), ("chain3", 777),
                                           ("body32", 5_009), ("body22", 7), ("body16", 1), ("deep65", 29)])
def test_fk_track_kernel(sk, set_knobs, knobs, name, n_frames):
    """The track kernel (U independent joints per step from the host's box-by-box level schedule, ring of TMA boxes,
    stage at the 16-byte phase of the global span with head / tail words stored by single lanes): every (U, NB) that
    is built, both tile sizes, odd joint counts with 10-frame tiles (unaligned spans), ragged frame counts (remainder
    tile), one warp per SM (stage and ring reuse across many tiles)."""
    set_knobs(knobs)
    par = parents_of(name)
    rot, gp, off = synth_numpy(n_frames, par, seed=13 * len(par) + n_frames)
    want_pos, want_rotm = orc.fk(rot, gp, off, par)
    for _ in range(2):
        pos, rotm = sk.fk(rot, gp, off, par)
        assert "fk_tracks_kernel" in _lib.load().pmb_last_variant().decode()
        assert_allclose(pos, want_pos, **TOL)
        assert_allclose(rotm, want_rotm, **TOL)


@pytest.mark.parametrize("knobs", [{"PMB_FK_MTRACKS": "1"}, {"PMB_FK_MTRACKS": "1", "PMB_FK_WARPS_PER_SM": "1"},
                                   {"PMB_FK_MTRACKS": "1", "PMB_FK_WARPS_PER_SM": "3"}])
@pytest.mark.parametrize("name,n_frames", [("body22", 40_003), ("smplh52", 20_051), ("deep65", 10_031), ("chain3", 777),
                                           ("body32", 5_009), ("body22", 7), ("body16", 1), ("deep65", 29), ("body40", 4_096)])
def test_fk_matrix_track_kernel(sk, set_knobs, knobs, name, n_frames):
    """The matrix track kernel (lanes = 4 tracks x 8 frames, whole 3x4 transform per lane, dense stage doubling as the parent
    store, bulk-stored full tiles, remainder tile copied out by the lanes): ragged frame counts, batches of less than a tile,
    one warp per SM (stage, input double buffer and barriers reused across many tiles)."""
    set_knobs(knobs)
    par = parents_of(name)
    rot, gp, off = synth_numpy(n_frames, par, seed=19 * len(par) + n_frames)
    want_pos, want_rotm = orc.fk(rot, gp, off, par)
    for _ in range(2):
        pos, rotm = sk.fk(rot, gp, off, par)
        assert "fk_mtracks_kernel" in _lib.load().pmb_last_variant().decode()
        assert_allclose(pos, want_pos, **TOL)
        assert_allclose(rotm, want_rotm, **TOL)


@pytest.mark.parametrize("seed", range(6))
def test_fk_track_kernel_random_trees(sk, set_knobs, seed):
    """Random topologies (bushy, deep, up to 200 joints) through the track schedule with 1 and 2 tracks."""
    rng = np.random.default_rng(1000 + seed)
    n_joints = int(rng.integers(2, 200))
    par = np.zeros(n_joints, dtype=np.int64)
    for i in range(1, n_joints):
        par[i] = rng.integers(max(0, i - 1 - int(rng.integers(0, 12))), i)
    n_frames = int(rng.integers(1, 3000))
    rot, gp, off = synth_numpy(n_frames, par, seed=seed)
    want_pos, want_rotm = orc.fk(rot, gp, off, par)
    for u, nb, ul in ((1, 2, 1), (2, 3, 1), (1, 3, 2), (2, 4, 2)):
        set_knobs({"PMB_FK_TRACKS": "1", "PMB_FK_U": str(u), "PMB_FK_NB": str(nb), "PMB_FK_UL": str(ul),
                   "PMB_FK_FR": "5" if ul == 2 else ("8" if n_joints > 150 else "10")})
        pos, rotm = sk.fk(rot, gp, off, par)
        assert "fk_tracks_kernel" in _lib.load().pmb_last_variant().decode()
        assert_allclose(pos, want_pos, rtol=2e-5, atol=2e-5)
        assert_allclose(rotm, want_rotm, rtol=2e-5, atol=2e-5)
    if n_joints <= 80:  # the matrix track kernel keeps a tile's whole output (384 J bytes per frame) in shared memory
        set_knobs({"PMB_FK_MTRACKS": "1"})
        pos, rotm = sk.fk(rot, gp, off, par)
        assert "fk_mtracks_kernel" in _lib.load().pmb_last_variant().decode()
        assert_allclose(pos, want_pos, rtol=2e-5, atol=2e-5)
        assert_allclose(rotm, want_rotm, rtol=2e-5, atol=2e-5)


def test_knobs_are_ignored_without_the_experiment_switch(sk, monkeypatch):
    """A stray PMB_* variable in a user's environment must not change which kernel runs (or make a call fail):
    knobs are honoured only under PMB_EXPERIMENT=1."""
    lib = _lib.load()
    par = parents_of("body22")
    rot, gp, off = synth_numpy(257, par, seed=5)
    monkeypatch.delenv("PMB_EXPERIMENT", raising=False)
    lib.pmb_reload_knobs()
    sk.fk(rot, gp, off, par)
    default_variant = lib.pmb_last_variant().decode()
    monkeypatch.setenv("PMB_FK_GROUP", "8")
    monkeypatch.setenv("PMB_FK_WARPS", "7")  # selects no variant when honoured
    lib.pmb_reload_knobs()
    try:
        sk.fk(rot, gp, off, par)
        assert lib.pmb_last_variant().decode() == default_variant
    finally:
        monkeypatch.undo()
        lib.pmb_reload_knobs()


@pytest.mark.parametrize("group", [None, "8", "16", "24"])
@pytest.mark.parametrize("name,n_frames", [("body22", 4099), ("smplh52", 2050), ("deep65", 1031), ("chain3", 777)])
def test_to_root_dual_quat_every_group(sk, set_knobs, group, name, n_frames):
    if group is not None:
        set_knobs({"PMB_DQ_GROUP": group})
    par = parents_of(name)
    rot, gp, off = synth_numpy(n_frames, par, seed=7 * len(par) + n_frames)
    dq = sk.to_root_dual_quat(rot, gp, par, off)
    assert_allclose(dq, orc.to_root_dual_quat(rot, gp, par, off), **TOL)


@pytest.mark.parametrize("knobs", [{"PMB_DQ_TRACKS": "1"}, {"PMB_DQ_TRACKS": "1", "PMB_QT_WARPS_PER_SM": "1"},
                                   {"PMB_DQ_TRACKS": "1", "PMB_QT_DYNAMIC": "0", "PMB_QT_WARPS_PER_SM": "3"}, {"PMB_DQ_TRACKS": "0"},
                                   {"PMB_DQ_TRACKS": "1", "PMB_QT_SHAPE": "1"}, {"PMB_DQ_TRACKS": "1", "PMB_QT_SHAPE": "2"},
                                   {"PMB_DQ_TRACKS": "1", "PMB_QT_SHAPE": "2", "PMB_QT_WARPS_PER_SM": "1"}])
@pytest.mark.parametrize("name,n_frames", [("body22", 40_003), ("smplh52", 20_051), ("deep65", 10_031), ("chain3", 777),
                                           ("body32", 5_009), ("body22", 7), ("body16", 1), ("deep65", 29)])
def test_to_root_dual_quat_track_kernel(sk, set_knobs, knobs, name, n_frames):
    """to_root_dual_quat through the quaternion track kernel (lanes = 4 tracks x 8 frames, whole-skeleton level schedule,
    dense output stage doubling as the parent store), forced on and off: ragged frame counts (remainder tile), one warp
    per SM (stage and input double buffer reused across many tiles), single-tile batches."""
    set_knobs(knobs)
    par = parents_of(name)
    rot, gp, off = synth_numpy(n_frames, par, seed=17 * len(par) + n_frames)
    want = orc.to_root_dual_quat(rot, gp, par, off)
    for _ in range(2):
        dq = sk.to_root_dual_quat(rot, gp, par, off)
        assert ("qtracks_kernel" in _lib.load().pmb_last_variant().decode()) == (knobs["PMB_DQ_TRACKS"] == "1")
        assert_allclose(dq, want, **TOL)


@pytest.mark.parametrize("seed", range(6))
def test_quaternion_track_kernel_random_trees(sk, set_knobs, seed):
    """Random topologies (bushy, deep, up to 200 joints) through the four-track whole-skeleton schedule: dual quaternions and
    fk_quat (positions at the hot-path tolerance, rotations as the same rotation with the reference's sign rule)."""
    rng = np.random.default_rng(2000 + seed)
    n_joints = int(rng.integers(2, 200))
    par = np.zeros(n_joints, dtype=np.int64)
    for i in range(1, n_joints):
        par[i] = rng.integers(max(0, i - 1 - int(rng.integers(0, 12))), i)
    n_frames = int(rng.integers(1, 3000))
    rot, gp, off = synth_numpy(n_frames, par, seed=seed)
    set_knobs({"PMB_DQ_TRACKS": "1", "PMB_FKQ_TRACKS": "1"})
    dq = sk.to_root_dual_quat(rot, gp, par, off)
    assert "qtracks_kernel" in _lib.load().pmb_last_variant().decode()
    assert_allclose(dq, orc.to_root_dual_quat(rot, gp, par, off), rtol=2e-5, atol=2e-5)
    want_pos, want_rotm = orc.fk(rot, gp, off, par)
    pos, grot = sk.fk_quat(rot, gp, off, par)
    assert "qtracks_kernel" in _lib.load().pmb_last_variant().decode()
    assert_allclose(pos, want_pos, rtol=2e-5, atol=2e-5)
    assert_allclose(np.abs(np.sum(grot * orc.quat_from_matrix(want_rotm), axis=-1)), 1.0, atol=2e-5)


def test_frame_shards_on_two_devices(sk):
    """Section 8e: the frame axis shards with no data-path collective.  Two shards computed on two GPUs of the box
    (one process, the library's per-device state) equal the single-GPU result bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from pymotion_b200.sharding import shard_bounds

    par = parents_of("body22")
    n = 100_003
    rot, gp, off = synth_numpy(n, par, seed=99)
    full_pos, full_rotm = sk.fk(torch.from_numpy(rot).cuda(0), torch.from_numpy(gp).cuda(0), torch.from_numpy(off).cuda(0), par)
    for rank in range(2):
        lo, hi = shard_bounds(n, 2, rank)
        dev = torch.device("cuda", rank)
        pos, rotm = sk.fk(torch.from_numpy(rot[lo:hi]).to(dev), torch.from_numpy(gp[lo:hi]).to(dev), torch.from_numpy(off).to(dev), par)
        assert pos.device == dev
        assert torch.equal(pos.cpu(), full_pos[lo:hi].cpu())
        assert torch.equal(rotm.cpu(), full_rotm[lo:hi].cpu())


@pytest.mark.parametrize("name,n_frames", [("body22", 1), ("body22", 1000), ("deep65", 2_049), ("body22", 120_001)])
def test_dual_quat_pair_on_host_arrays(sk, name, n_frames):
    """to_root_dual_quat / from_root_dual_quat called the reference's way -- NumPy in, NumPy out -- take the host pipeline
    (pmb_*_f32_host: chunked copies in / kernel / copies out on two streams; 120 001 x 22 is several chunks with a ragged last
    one) and give the bits of the device-resident path; float64 in -> float64 out; CPU tensors in -> CPU tensors out; the
    reference's offsets[0] assert (skeleton.py:227) survives; inputs are not modified."""
    par = parents_of(name)
    rot, gp, off = synth_numpy(n_frames, par, seed=23 * len(par) + n_frames)
    keep = [a.copy() for a in (rot, gp, off)]
    dq = sk.to_root_dual_quat(rot, gp, par, off)
    assert isinstance(dq, np.ndarray) and dq.dtype == np.float32 and dq.shape == (n_frames, len(par), 8)
    dev = [torch.from_numpy(a).cuda() for a in (rot, gp, off)]
    want = sk.to_root_dual_quat(dev[0], dev[1], par, dev[2])
    assert np.array_equal(dq, want.cpu().numpy())
    trans, rots = sk.from_root_dual_quat(dq, par)
    wt, wr = sk.from_root_dual_quat(want, par)
    assert isinstance(trans, np.ndarray) and np.array_equal(trans, wt.cpu().numpy()) and np.array_equal(rots, wr.cpu().numpy())
    for a, b in zip((rot, gp, off), keep):
        assert np.array_equal(a, b)
    if n_frames <= 2_049:
        assert_allclose(dq, orc.to_root_dual_quat(rot, gp, par, off), **TOL)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            d64 = sk.to_root_dual_quat(rot.astype(np.float64), gp.astype(np.float64), par, off.astype(np.float64))
        assert d64.dtype == np.float64
        assert_allclose(d64, dq, rtol=0, atol=0)
        t_cpu = sk.to_root_dual_quat(torch.from_numpy(rot), torch.from_numpy(gp), par, torch.from_numpy(off))
        assert isinstance(t_cpu, torch.Tensor) and not t_cpu.is_cuda and torch.equal(t_cpu, torch.from_numpy(dq))
        bad = off.copy()
        bad[0, 1] = 0.25
        with pytest.raises(AssertionError):
            sk.to_root_dual_quat(rot, gp, par, bad)
