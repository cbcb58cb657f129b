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
    res = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1  # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "skeleton poses/sec (22 joints)" and d["unit"] == "poses/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "fk_1m_x_22"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "poses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


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
