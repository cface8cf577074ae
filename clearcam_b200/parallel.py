"""Multi-GPU host logic: one process per GPU (torchrun), frames / crops / queries sharded by batch.

The detector needs no collective (results are per frame and stay with the owning rank); the CLIP
index does one exchange: an in-place all-gather of the [B_local, D] fp32 embeddings that the final kernel has
already written into this rank's slice of the gather buffer (NCCL over NVLink on GPU, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous near-equal split of n units: rank r owns [lo, hi)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(full: torch.Tensor, rows_per_rank: int, group=None) -> torch.Tensor:
    """In-place all-gather: `full` is [world*rows_per_rank, D]; this rank has already filled its own slice."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    mine = full[rank * rows_per_rank:(rank + 1) * rows_per_rank]
    dist.all_gather_into_tensor(full, mine, group=group)
    return full


def max_over_ranks(value: float, device, group=None) -> float:
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
