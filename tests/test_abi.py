"""CPU-only checks of the C-ABI boundary: the library builds/loads, exports every
symbol include/pymotion_b200.h declares, and its host-side logic (validation,
joint-program compiler) behaves -- no kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from pymotion_b200 import _build, _lib
from pymotion_b200.topologies import TOPOLOGIES, parents_of

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if _build.is_stale() and os.path.exists(_build.NVCC):
        _build.build()
    return _lib.load()


def header_functions():
    text = open(os.path.join(REPO, "include", "pymotion_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pmb_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    names = header_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/pymotion_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in pymotion_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.pmb_version() == 100


def test_built_for_sm_100a():
    out = os.popen(f"cuobjdump -lelf {_lib.LIB_PATH} 2>/dev/null").read()
    if not out:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out


def python_joint_program(parents):
    """Independent restatement of csrc/joint_program.h (linear-scan slot allocation)."""
    n = len(parents)
    last = [-1] * n
    for i in range(1, n):
        if parents[i] != i - 1:
            last[parents[i]] = i
    free, nxt, slot, codes = [], 0, [-1] * n, []
    for i in range(n):
        p = 0 if i == 0 else int(parents[i])
        src = save = 0xFF
        if i > 0 and p != i - 1:
            src = slot[p]
            if last[p] == i:
                free.append(slot[p])
        if last[i] >= 0:
            if free:
                s = min(free)
                free.remove(s)
            else:
                s, nxt = nxt, nxt + 1
            slot[i] = save = s
        codes.append(src | (save << 8) | (p << 16))
    return nxt, codes


@pytest.mark.parametrize("name", sorted(TOPOLOGIES))
def test_joint_program_matches_restatement(lib, name):
    par = parents_of(name)
    codes = np.zeros(len(par), dtype=np.uint32)
    n_slots = lib.pmb_build_joint_program(par.ctypes.data, len(par), codes.ctypes.data)
    want_slots, want_codes = python_joint_program(par)
    assert n_slots == want_slots
    assert codes.tolist() == want_codes
    # simulate the program: every fetch must return the transform of the right parent
    held = {}
    for i, c in enumerate(codes.tolist()):
        src, save, p = c & 0xFF, (c >> 8) & 0xFF, (c >> 16) & 0x7FFF
        if i > 0:
            assert p == par[i]
            if src == 0xFF:
                assert p == i - 1
            else:
                assert held[src] == p
        if save != 0xFF:
            held[save] = i
    assert {"body22": 1, "chain3": 0}.get(name, n_slots) == n_slots


def test_random_trees_program_is_consistent(lib):
    rng = np.random.default_rng(7)
    for n in (1, 2, 5, 33, 64, 200, 512):
        par = np.zeros(n, dtype=np.int64)
        for i in range(1, n):
            par[i] = rng.integers(0, i)
        codes = np.zeros(n, dtype=np.uint32)
        n_slots = lib.pmb_build_joint_program(par.ctypes.data, n, codes.ctypes.data)
        want_slots, want_codes = python_joint_program(par)
        assert (n_slots, codes.tolist()) == (want_slots, want_codes)


def test_topology_and_shape_errors_need_no_gpu(lib):
    dummy = np.zeros(64, dtype=np.float32)  # never dereferenced: validation fails first
    p = dummy.ctypes.data
    bad = np.array([0, 2, 0], dtype=np.int64)  # parents[1] = 2 >= 1
    rc = lib.pmb_fk_f32(p, p, 3, p, 0, bad.ctypes.data, 4, 3, p, p, None)
    assert rc == _lib.PMB_ERR_TOPOLOGY and "parents[1]" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    neg = np.array([0, -1, 0], dtype=np.int64)
    assert lib.pmb_fk_f32(p, p, 3, p, 0, neg.ctypes.data, 4, 3, p, p, None) == _lib.PMB_ERR_TOPOLOGY
    ok = np.array([-1, 0, 1], dtype=np.int64)  # parents[0] is ignored, like the reference
    assert lib.pmb_fk_f32(p, p, 3, p, 0, ok.ctypes.data, 0, 3, p, p, None) == _lib.PMB_OK  # 0 frames: no launch
    assert lib.pmb_fk_f32(None, p, 3, p, 0, ok.ctypes.data, 4, 3, p, p, None) == _lib.PMB_ERR_NULL
    assert lib.pmb_fk_f32(p, p, 2, p, 0, ok.ctypes.data, 4, 3, p, p, None) == _lib.PMB_ERR_SHAPE
    assert lib.pmb_fk_f32(p, p, 3, p, 0, ok.ctypes.data, 4, 0, p, p, None) == _lib.PMB_ERR_SHAPE
    assert lib.pmb_fk_f32(p + 4, p, 3, p, 0, ok.ctypes.data, 4, 3, p, p, None) == _lib.PMB_ERR_ALIGN
    off0 = np.array([0, 0.5, 0], dtype=np.float32)
    rc = lib.pmb_to_root_dual_quat_f32(p, p, 3, ok.ctypes.data, p, off0.ctypes.data, 4, 3, p, None)
    assert rc == _lib.PMB_ERR_ROOT_OFFSET
    with pytest.raises(AssertionError):  # ops/skeleton.py:227 is an assert
        _lib.check(rc)
    assert lib.pmb_status_string(rc) == b"offsets[0] != 0"
    assert lib.pmb_quat_mul_f32(p, p, p, -1, None) == _lib.PMB_ERR_SHAPE
    assert lib.pmb_quat_mul_f32(p, p, p, 0, None) == _lib.PMB_OK


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under pymotion_b200/ may import it."""
    pkg = os.path.join(REPO, "pymotion_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "oracle" not in text.lower() or f == "_lib.py" and False, f"{f} mentions the oracle"


def _track_schedule(lib, parents, n_tracks, window=0):
    par = np.asarray(parents, dtype=np.int64)
    codes = np.zeros(1024, dtype=np.uint32)
    first = np.zeros(70, dtype=np.int32)
    steps = lib.pmb_build_track_schedule(par.ctypes.data, len(par), n_tracks, window, codes.ctypes.data, codes.size, first.ctypes.data)
    assert steps > 0, _lib.last_error()
    n_windows = -(-len(par) // window) if window else 1
    return steps, codes[: steps * n_tracks].reshape(steps, n_tracks), first[: n_windows + 1]


def _check_schedule(par, code, n_tracks):
    """Every joint exactly once, after its parent's step; `carry` only when the parent is the same track's
    previous item.  Returns step_of."""
    CARRY, NOOP = 1 << 20, 1 << 21
    step_of, track_of = {}, {}
    for t in range(code.shape[0]):
        for u in range(n_tracks):
            c = int(code[t, u])
            if c & NOOP:
                continue
            j, p = c & 0x3FF, (c >> 10) & 0x3FF
            assert j not in step_of, f"joint {j} scheduled twice"
            step_of[j], track_of[j] = t, u
            if j == 0:
                assert t == 0 and u == 0 and (c & CARRY)
                continue
            assert p == par[j]
            assert step_of[p] < t, f"joint {j} at step {t} before its parent {p}"
            if c & CARRY:
                assert step_of[p] == t - 1 and track_of[p] == u
    assert sorted(step_of) == list(range(len(par)))
    return step_of


@pytest.mark.parametrize("name", ["body22", "smplh52", "deep65", "chain3", "body32"])
@pytest.mark.parametrize("n_tracks", [1, 2, 3, 4])
def test_track_schedule_is_a_valid_minimal_tree_schedule(name, n_tracks):
    """The whole-skeleton schedule is the optimum for unit tasks with tree precedence on `n_tracks` machines
    (Hu's bound: max over levels l of  l + 1 + ceil(#joints deeper than l / n_tracks))."""
    lib = _lib.load()
    par = parents_of(name)
    n = len(par)
    steps, code, first = _track_schedule(lib, par, n_tracks)
    _check_schedule(par, code, n_tracks)
    assert list(first) == [0, steps]
    depth = np.zeros(n, dtype=int)
    for i in range(1, n):
        depth[i] = depth[par[i]] + 1
    bound = max(l + 1 + -(-int((depth > l).sum()) // n_tracks) for l in range(int(depth.max()) + 1))
    assert steps == max(bound, int(depth.max()) + 1)


@pytest.mark.parametrize("name", sorted(TOPOLOGIES))
@pytest.mark.parametrize("n_tracks", [1, 2, 3])
def test_track_schedule_by_windows_of_one_box(name, n_tracks):
    """What the fk track kernel runs: joints scheduled box by box (8 consecutive joints).  Valid as a whole, every
    step holds joints of ONE window only, windows come in order, and the window table brackets them."""
    lib = _lib.load()
    par = parents_of(name)
    n = len(par)
    steps, code, first = _track_schedule(lib, par, n_tracks, window=8)
    step_of = _check_schedule(par, code, n_tracks)
    assert first[0] == 0 and first[-1] == steps and all(a < b for a, b in zip(first, first[1:]))
    for j in range(n):
        w = j // 8
        assert first[w] <= step_of[j] < first[w + 1], f"joint {j} outside the steps of its window"
    if n_tracks == 1:
        assert steps == n
    assert steps <= n and steps >= -(-n // n_tracks)


def test_track_schedule_rejects_bad_tables():
    lib = _lib.load()
    codes = np.zeros(64, dtype=np.uint32)
    bad = np.array([0, 2, 1], dtype=np.int64)
    assert lib.pmb_build_track_schedule(bad.ctypes.data, 3, 2, 0, codes.ctypes.data, 64, None) == _lib.PMB_ERR_TOPOLOGY
    ok = np.array([0, 0, 1], dtype=np.int64)
    assert lib.pmb_build_track_schedule(ok.ctypes.data, 3, 9, 0, codes.ctypes.data, 64, None) == _lib.PMB_ERR_SHAPE
    assert lib.pmb_build_track_schedule(ok.ctypes.data, 3, 2, 0, codes.ctypes.data, 2, None) == _lib.PMB_ERR_SHAPE
    assert lib.pmb_build_track_schedule(ok.ctypes.data, 3, 2, 5, codes.ctypes.data, 64, None) == _lib.PMB_ERR_SHAPE
