"""Multi-GPU plumbing for the hot path: it shards by sequence (independent units), one process per
GPU, and has NO data-path collective; torch.distributed is only used to reduce timings (max over ranks)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of n_items sequences: ranks < n % world get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sequences(items: Sequence, rank: int, world: int) -> List:
    a, b = shard_range(len(items), rank, world)
    return list(items[a:b])


def max_over_ranks(values: Sequence[float], device: torch.device) -> List[float]:
    """Element-wise maximum of per-rank measurements (device timings are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()
