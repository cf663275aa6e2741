"""Multi-GPU sharding of the dense grid: one process per GPU, contiguous ranges
of the flattened res^3 grid per rank (x-slabs), ONE all_gather of fp32 SDF values at the end.
The reference's only parallelism is nn.DataParallel (train.py:126, test.py:62), a no-op at test
time; nothing in rows a-2..a-6 mixes query points, so the partition is exact."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int, align: int = 1) -> Tuple[int, int]:
    """[begin, count) of `rank`: equal contiguous ranges, boundaries rounded to `align`
    (the last rank takes the remainder)."""
    per = -(-total // world)
    per = -(-per // align) * align
    begin = min(total, rank * per)
    return begin, max(0, min(total, begin + per) - begin)


def gather_shards(local: torch.Tensor, total: int, world: int, align: int = 1, group=None) -> torch.Tensor:
    """local: (B, count_r) fp32 shard of every image -> (B, total) on every rank, one collective."""
    if world == 1:
        return local
    B = local.shape[0]
    per = shard_range(total, 0, world, align)[1]
    buf = torch.zeros(B, per, device=local.device, dtype=local.dtype)
    buf[:, :local.shape[1]] = local
    out = torch.empty(world * B, per, device=local.device, dtype=local.dtype)      # rank-major concat on dim 0
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    return out.view(world, B, per).permute(1, 0, 2).reshape(B, world * per)[:, :total].contiguous()


def sharded_grid(evaluate: Callable[[int, int], torch.Tensor], total: int, align: int = 1, group=None) -> torch.Tensor:
    """evaluate(begin, count) -> (B, count); returns the gathered (B, total) grid on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return evaluate(0, total)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    begin, count = shard_range(total, rank, world, align)
    return gather_shards(evaluate(begin, count), total, world, align, group)
