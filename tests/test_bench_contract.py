"""bench.py on a box without a GPU: the reference arm prints the contract line, our arm refuses to run (there is no
CPU fallback), and the byte accounting matches SURVEY.md section 8d."""
import json
import os
import subprocess
import sys

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    env = {k: v for k, v in os.environ.items() if not k.startswith("PMB_")}
    return subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), *args], capture_output=True, text=True,
                          cwd=REPO, env=env, timeout=600)


def test_reference_arm_contract_line():
    res = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--ref-frames", "40000")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1  # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "skeleton poses/sec (22 joints)" and d["unit"] == "poses/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None
    # the SAME config object as our arm prints (bench.shared_config): the driver compares the two
    sys.path.insert(0, REPO)
    import bench

    assert d["config"] == bench.shared_config("fk_1m_x_22", 1)
    # the unmodified reference when it has been staged under oracle/_ref (build()), else the NumPy port
    staged = os.path.isdir(os.path.join(REPO, "oracle", "_ref", "pymotion"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_staged_reference_is_the_reference_and_agrees_with_the_oracle():
    """oracle/_ref (staged by oracle/fetch_ref.py at build()) is what `--impl reference` times; the oracle port the
    tests use must agree with it on the bench's own inputs."""
    import numpy as np

    from oracle import fetch_ref
    from oracle import pymotion_oracle as orc
    from pymotion_b200.topologies import parents_of, synth_numpy

    mods = fetch_ref.import_reference()
    if mods is None:
        pytest.skip("oracle/_ref has not been staged (no /root/reference on this box)")
    sk = mods[0]
    assert os.path.realpath(sk.__file__).startswith(os.path.realpath(os.path.join(REPO, "oracle", "_ref")))
    par = parents_of("body22")
    rot, gpos, off = synth_numpy(500, par, seed=0)
    pos, rotm = sk.fk(rot, gpos, off, par)
    want_pos, want_rotm = orc.fk(rot, gpos, off, par)
    np.testing.assert_allclose(pos, want_pos, rtol=0, atol=1e-12)
    np.testing.assert_allclose(rotm, want_rotm, rtol=0, atol=1e-12)


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_our_arm_fails_loudly_without_a_gpu():
    res = run_bench("--steps", "1", "--warmup", "1")
    assert res.returncode != 0
    assert "{" not in res.stdout  # no line, no number
    assert "CUDA" in res.stderr or "cuda" in res.stderr


def test_algorithmic_bytes_per_pose():
    sys.path.insert(0, REPO)
    import bench

    assert bench.fk_bytes_per_pose(22) == 1420 and bench.fk_bytes_per_pose(52) == 3340 and bench.fk_bytes_per_pose(65) == 4172
    assert bench.op_bytes_per_pose("to_dq", 22) == 1068 and bench.op_bytes_per_pose("from_dq", 22) == 1320
    assert bench.op_bytes_per_pose("round_trip", 22) == 2388
    assert bench.metric_of(22) == bench.METRIC
    assert set(bench.WORKLOADS) >= {"fk_1m_x_22", "fk_4m_x_52", "fk_4m_x_65"}
