"""Multi-GPU plumbing.  The forward path shards by image (one process per GPU, weights
replicated, NO collective); the only exchange is an optional all-gather of the padded
detections.  Backend-agnostic torch.distributed calls so the host logic is
testable with gloo on CPU."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of `total` images over `world` ranks (first ranks take the remainder)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_detections(dets: torch.Tensor, counts: torch.Tensor):
    """dets [B_local, max_det, 6] f32, counts [B_local] i32 -> ([B_total, max_det, 6], [B_total]),
    ranks concatenated in rank order.  Every rank must pass the same shapes."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return dets, counts
    d = [torch.empty_like(dets) for _ in range(world)]
    c = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(d, dets.contiguous())
    dist.all_gather(c, counts.contiguous())
    return torch.cat(d, 0), torch.cat(c, 0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
