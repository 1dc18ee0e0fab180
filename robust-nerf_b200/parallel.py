"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

Training: rays are independent, so the batch is sharded and the ONLY exchange is one all-reduce
(SUM, then 1/world) of the flat gradient buffer [grad coarse | grad fine | grad omega | grad delta_t]
(4,766,752 + 2,400 bytes) per step -- NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests.
Rendering: tiles of rays are dealt round-robin to ranks; no collective.
These helpers are device-agnostic so the N>1 logic is testable with gloo on CPU.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def allreduce_mean_(flat: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """In-place mean over ranks of a flat buffer (equal shard sizes => exact global-batch mean)."""
    if world > 1:
        torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)
    return flat


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous equal shard [a, b) of a global batch of n items (n must divide evenly)."""
    if n % world != 0:
        raise ValueError(f"global batch {n} is not divisible by world size {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def tiles_for_rank(view: int, tiles_per_view: int, rank: int, world: int) -> List[int]:
    """Tiles k of `view` owned by `rank`: global tile id view*tiles_per_view + k, dealt round-robin."""
    return [k for k in range(tiles_per_view) if (view * tiles_per_view + k) % world == rank]
