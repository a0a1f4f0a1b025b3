"""Row-range sharding for the one-process-per-GPU launch.

The filter/project/limit path has no cross-partition reduction: GPU g of G owns the contiguous rows
[g*ceil(N/G), (g+1)*ceil(N/G)) (boundaries rounded to 64 rows so bitmap words never straddle GPUs), runs the
fused kernel locally, and the ordered result is the concatenation of the per-GPU outputs in rank order.  The
only cross-rank datum is each shard's survivor count (G integers), exchanged through torch.distributed
(NCCL on GPUs, gloo in the CPU tests) — there is no data-path collective.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import capi


def shard_rows(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of `rank`'s shard (rvl_shard_range)."""
    return capi.shard_range(n_rows, rank, world)


def ordered_offsets(counts: Sequence[int]) -> List[int]:
    """Exclusive scan of the per-shard survivor counts: where each shard's output starts in the ordered result."""
    out, run = [], 0
    for c in counts:
        out.append(run)
        run += int(c)
    return out


def limit_take(counts: Sequence[int], limit: int) -> List[int]:
    """Rows shard g contributes under a global LIMIT: clamp(limit - sum_{j<g} counts[j], 0, counts[g])."""
    return capi.shard_limit_split([int(c) for c in counts], limit)


def exchange_counts(local_count: int, device=None) -> List[int]:
    """All-gather one integer per rank (the only cross-GPU exchange of the path)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(local_count)]
    world = dist.get_world_size()
    t = torch.tensor([int(local_count)], dtype=torch.int64, device=device if device is not None else "cpu")
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    return [int(g.item()) for g in gathered]
