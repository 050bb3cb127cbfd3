"""Multi-GPU partitioning of the hot path: chunks are independent, so rows of the
``[N,1,L]`` batch are split into contiguous, near-equal slices, one per rank; no collective inside the loop.
The only communication is the final gather of the enhanced rows (NCCL on GPUs, gloo in CPU tests)."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of rank `rank`; the first n_rows % world ranks get one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank %d/%d" % (rank, world))
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_bounds(n_rows: int, world: int) -> List[Tuple[int, int]]:
    return [shard_bounds(n_rows, world, r) for r in range(world)]


def gather_rows(local: torch.Tensor, n_rows: int, group=None) -> torch.Tensor:
    """All-gather the per-rank slices back into [n_rows, ...] (every rank gets the full result).

    Slices may differ by one row, so every rank pads to the maximum slice before the collective."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world = dist.get_world_size(group)
    bounds = all_bounds(n_rows, world)
    width = max(hi - lo for lo, hi in bounds)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, bounds)], dim=0)
