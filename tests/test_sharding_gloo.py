"""N > 1 host logic on CPU: frame-axis partition and the optional gather of positions, over gloo with
world size 2 (two real processes).  The per-shard compute is stood in for by the oracle -- this test is
about the sharding, not the kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pymotion_b200 import sharding


def test_shard_bounds_tile_the_frame_axis():
    for n in (0, 1, 7, 8, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            blocks = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pymotion_oracle as orc
        from pymotion_b200.topologies import parents_of, synth_numpy

        par = parents_of("body22")
        rot, gp, off = synth_numpy(n_frames, par, seed=5)  # same full batch on every rank
        lo, hi = sharding.shard_bounds(n_frames, world, rank)
        assert sharding.shard_frames(torch.from_numpy(rot)).shape[0] == hi - lo
        pos_local, _ = orc.fk(rot[lo:hi], gp[lo:hi], off, par)  # no collective in the compute
        full = sharding.all_gather_frames(torch.from_numpy(np.ascontiguousarray(pos_local)), n_frames)
        want, _ = orc.fk(rot, gp, off, par)
        np.testing.assert_allclose(full.numpy(), want, rtol=0, atol=0)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == world
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [64, 101])
def test_two_rank_gather_matches_single_process(tmp_path, n_frames):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
