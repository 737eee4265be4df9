"""Multi-GPU partitioning of the hot path: independent clips, one process per GPU.

Every clip (batch row) is independent end to end (speinet.py:150-168 only masks rows), so the path
shards with NO data-path collective: rank r owns the clips with `clip_id % world == r`
(SURVEY.md section 8(e)).  The only exchange is one all-gather of the per-clip outputs at the end,
replacing the reference's single-process `nn.DataParallel` gather (inference_SPEINet.py:235,569).
Works with NCCL (GPU tensors) and gloo (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def shard_clips(num_clips: int, rank: int, world: int) -> List[int]:
    """Clip ids owned by `rank` (round-robin, so ragged totals differ by at most one clip)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, num_clips, world))


def gather_outputs(local: torch.Tensor, num_clips: int, rank: int, world: int, group=None) -> torch.Tensor:
    """All-gather per-clip outputs back into clip order.

    `local` is [len(shard_clips(num_clips, rank, world)), ...]; the result is [num_clips, ...] on every
    rank.  Ragged shards are padded to the longest shard for the collective and trimmed afterwards."""
    mine = shard_clips(num_clips, rank, world)
    if local.shape[0] != len(mine):
        raise ValueError(f"rank {rank} holds {local.shape[0]} clips, expected {len(mine)}")
    if world == 1:
        return local
    per = (num_clips + world - 1) // world
    tail = local.shape[1:]
    send = local.new_zeros((per,) + tuple(tail))
    send[:len(mine)] = local
    recv = local.new_empty((world * per,) + tuple(tail))
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    out = local.new_empty((num_clips,) + tuple(tail))
    recv = recv.view((world, per) + tuple(tail))
    for r in range(world):
        ids = shard_clips(num_clips, r, world)
        if ids:
            out[ids] = recv[r, :len(ids)]
    return out
