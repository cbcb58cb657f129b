"""Frame-axis sharding across the GPUs of one box.

Every frame is independent in fk / to_root_dual_quat / from_root_dual_quat, so the
batch splits into contiguous frame blocks, one per rank (one process per GPU), and
the compute needs NO collective.  The only exchange the path ever wants is the
optional final gather of global positions, done here with torch.distributed
(NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_frames: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous block of rank `rank`: sizes differ by at most one frame, blocks tile [0, n_frames)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    base, extra = divmod(int(n_frames), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_frames(x: torch.Tensor, world_size: int | None = None, rank: int | None = None) -> torch.Tensor:
    """This rank's frames of a [n_frames, ...] array (a view, no copy)."""
    world_size = dist.get_world_size() if world_size is None else world_size
    rank = dist.get_rank() if rank is None else rank
    lo, hi = shard_bounds(x.shape[0], world_size, rank)
    return x[lo:hi]


def all_gather_frames(local: torch.Tensor, n_frames: int, group=None) -> torch.Tensor:
    """Gather every rank's [shard, ...] block into the full [n_frames, ...] array on every rank.
    Shards may differ by one frame, so blocks are padded to the largest shard for the collective."""
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (shard_bounds(n_frames, world, r) for r in range(world))]
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError(f"local shard has {local.shape[0]} frames, expected {sizes[dist.get_rank(group)]}")
    biggest = max(sizes)
    padded = local
    if local.shape[0] != biggest:
        padded = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
    out = torch.empty((world, biggest) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if local.is_cuda:  # NCCL writes every block straight into its place of the one output buffer
        dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    else:
        dist.all_gather(list(out.unbind(0)), padded.contiguous(), group=group)  # gloo (CPU tests)
    if min(sizes) == biggest:
        return out.view((world * biggest,) + tuple(local.shape[1:]))  # equal shards: the buffer IS the result
    return torch.cat([out[r, : sizes[r]] for r in range(world)], dim=0)


def fk_sharded(rot, global_pos, offsets, parents, gather_positions: bool = False, group=None):
    """fk on this rank's frame block of the FULL [n_frames, ...] inputs.

    Returns (positions, rotmats) of the local block; with `gather_positions` the positions of every rank
    are all-gathered (the optional exchange of BASELINE.json config 5) and returned in full."""
    from .ops import skeleton

    rot, global_pos, offsets = torch.as_tensor(rot), torch.as_tensor(global_pos), torch.as_tensor(offsets)
    if rot.dim() != 3:
        raise ValueError(f"fk_sharded shards the leading axis of rot [n_frames, n_joints, 4], got {tuple(rot.shape)}")
    n_frames = rot.shape[0]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_bounds(n_frames, world, rank)
    # per-frame operands are recognised by RANK (a shared [3] root position must not be sliced when n_frames == 3)
    gp = global_pos[lo:hi] if (global_pos.dim() == rot.dim() - 1 and global_pos.shape[0] == n_frames) else global_pos
    off = offsets[lo:hi] if offsets.dim() == 3 else offsets
    pos, rotm = skeleton.fk(rot[lo:hi], gp, off, parents)
    if gather_positions:
        pos = all_gather_frames(pos, n_frames, group)
    return pos, rotm
