"""Frame sharding across the GPUs of one box.

Frames share no state on this path (the key space is partitioned by frame index,
dynamic_pillar_vfe.py:104; canvases are per frame, pointpillar_scatter.py:18-32), so a batch is split
into contiguous blocks of frames, one block per rank, with NO collective on the hot path.  A frame is
never split across ranks.  The only collective is an optional all-gather of the per-rank BEV blocks for
validation, outside any timed region.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def frame_range(num_frames: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [first, last) of frames owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(int(num_frames), int(world_size))
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def shard_points(points: torch.Tensor, num_frames: int, rank: int, world_size: int) -> Tuple[torch.Tensor, int]:
    """Rows of ``points`` (N, 1+C; column 0 = frame index) whose frame belongs to ``rank``, with the frame
    index renumbered from 0.  Returns (local_points, local_num_frames)."""
    first, last = frame_range(num_frames, rank, world_size)
    b = points[:, 0]
    keep = (b >= first) & (b < last)
    local = points[keep].clone()
    local[:, 0] -= first
    return local.contiguous(), last - first


def gather_bev(local_bev: torch.Tensor, num_frames: int, group=None) -> torch.Tensor:
    """All-gather the per-rank (B_local, C, ny, nx) blocks into the full (num_frames, C, ny, nx) map on
    every rank (validation only).  Blocks may differ by one frame, so they are padded to the largest."""
    world = dist.get_world_size(group)
    if world == 1:
        return local_bev
    sizes = [frame_range(num_frames, r, world) for r in range(world)]
    bmax = max(l - f for f, l in sizes)
    pad = local_bev
    if local_bev.shape[0] < bmax:
        pad = torch.zeros((bmax,) + tuple(local_bev.shape[1:]), dtype=local_bev.dtype, device=local_bev.device)
        pad[:local_bev.shape[0]] = local_bev
    out = torch.empty((world * bmax,) + tuple(local_bev.shape[1:]), dtype=local_bev.dtype, device=local_bev.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    blocks: List[torch.Tensor] = []
    for r, (f, l) in enumerate(sizes):
        blocks.append(out[r * bmax: r * bmax + (l - f)])
    return torch.cat(blocks, dim=0)
