"""Multi-GPU plumbing for the hot path: one process per GPU, replicas only.

Videos are independent on the inference path (no cross-sample op; BatchNorm in eval mode), so the
path shards by clips with NO data-path collective.  These helpers are the only
places where ranks talk: splitting a batch, gathering the (B,1) logits when a caller wants them in
one place, and the max-over-ranks used for timing.  They work on any torch.distributed backend
(NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, stop) of `total` clips for `rank` (first `total % world` ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def max_over_ranks(value: float, device=None) -> float:
    """max over ranks of a per-rank scalar (device-timed milliseconds in bench.py)."""
    dist = _dist()
    if dist is None:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_rows(t: torch.Tensor, total: int) -> torch.Tensor:
    """Concatenate per-rank row shards (made with shard_range) back into rank order: (total, ...)."""
    dist = _dist()
    if dist is None:
        return t
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world) for r in range(world)]
    maxn = max(b - a for a, b in sizes)
    pad = torch.zeros((maxn,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)


def barrier() -> None:
    dist = _dist()
    if dist is not None:
        dist.barrier()
