"""Multi-GPU plumbing for a path that needs no data-path collective.

Utterances (and streaming sessions) are independent in eval mode, so the batch is cut into contiguous per-rank
slices.  ``torch.distributed`` is used only to agree on timing: a barrier around the timed region and
a MAX reduction of the per-rank device time.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def shard_slice(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of ``total`` utterances owned by ``rank``; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """MAX of a per-rank scalar over the default process group (identity when not initialised)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(device: Optional[torch.device] = None) -> None:
    import torch.distributed as dist

    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)
