"""Row-range sharding for the one-process-per-GPU launch.

The filter/project/limit path has no cross-partition reduction: GPU g of G owns the contiguous rows
[g*ceil(N/G), (g+1)*ceil(N/G)) (boundaries rounded to 64 rows so bitmap words never straddle GPUs), runs the
operator locally, and the ordered result is the concatenation of the per-GPU outputs in rank order
(RecordBatch::concat, /root/reference/src/execution/record_batch.rs:245-342; collect_stream_batches,
physical_plan/streaming.rs:343-352).  The only cross-rank DATA exchange is optional: `gather_ordered` lands that
concatenation physically on one GPU, every rank writing its own rows into the destination's memory over NVLink at
once (CUDA IPC, rvl_gather_*).  What always crosses ranks is G integers per column — the survivor counts (and string
byte counts) whose exclusive scan gives every shard its place — exchanged through torch.distributed (NCCL on GPUs,
gloo in the CPU tests).  No NCCL collective touches a data buffer.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

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


def exchange_ints(values: Sequence[int], device=None) -> List[List[int]]:
    """All-gather a short vector of integers per rank: result[r] is rank r's vector (the path's only mandatory exchange)."""
    import torch
    import torch.distributed as dist

    vals = [int(v) for v in values]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [vals]
    world = dist.get_world_size()
    t = torch.tensor(vals, dtype=torch.int64, device=device if device is not None else "cpu")
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    return [[int(x) for x in g.tolist()] for g in gathered]


def exchange_counts(local_count: int, device=None) -> List[int]:
    """All-gather one integer per rank."""
    return [v[0] for v in exchange_ints([local_count], device)]


def plan_gather(per_rank: Sequence[Sequence[int]], limit: int = -1):
    """From every rank's [rows, string bytes of column 0, 1, ...] vector: (rows each shard contributes under `limit`, row offset
    of each shard, per-shard byte offsets per column, total rows, total bytes per column).  Pure arithmetic — the CPU tests
    check it against a sequential concatenation."""
    counts = [int(v[0]) for v in per_rank]
    take = limit_take(counts, limit) if limit >= 0 else counts
    row_off = ordered_offsets(take)
    ncol = len(per_rank[0]) - 1
    byte_off = [[0] * ncol for _ in per_rank]
    totals = [0] * ncol
    for c in range(ncol):
        run = 0
        for r, v in enumerate(per_rank):
            byte_off[r][c] = run
            run += int(v[1 + c])
        totals[c] = run
    return take, row_off, byte_off, sum(take), totals


def gather_ordered(ctx: "capi.Context", part: "capi.Batch", dst_rank: int = 0, device=None, timer=None) -> Optional["capi.Batch"]:
    """Order-preserving physical concatenation of every rank's `part` on rank `dst_rank`'s GPU, one process per GPU.

    counts -> offsets (all-gather of a few integers), destination buffers exported over CUDA IPC and broadcast as bytes,
    then every rank pushes its rows into the destination's memory concurrently (NVLink peer writes).  Returns the gathered
    batch on `dst_rank`, None elsewhere.  `timer(label)` is called at phase boundaries (bench.py times the push phase)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    ncols = part.num_columns()
    views = [part.view(j) for j in range(ncols)]
    dtypes = [v.dtype for v in views]
    str_cols = [j for j in range(ncols) if dtypes[j] == capi.STRING]
    # bytes referenced by the part's string columns (outputs of the operator start at byte 0 and are dense: data_len is exact)
    mine = [part.num_rows()] + [int(views[j].data_len) for j in str_cols] + [int(views[j].validity is not None) for j in range(ncols)]
    allv = exchange_ints(mine, device)
    per_rank = [v[:1 + len(str_cols)] for v in allv]
    take, row_off, byte_off, total_rows, total_bytes = plan_gather(per_rank)
    has_validity = [any(v[1 + len(str_cols) + j] for v in allv) for j in range(ncols)]
    data_bytes = [0] * ncols
    my_byte_off = [0] * ncols
    for i, j in enumerate(str_cols):
        data_bytes[j] = total_bytes[i]
        my_byte_off[j] = byte_off[rank][i]
    dest = None
    if rank == dst_rank:
        dest, blob = capi.gather_dest_create(ctx, dtypes, has_validity, total_rows, data_bytes)
        payload = [blob]
    else:
        payload = [None]
    if world > 1:
        dist.broadcast_object_list(payload, src=dst_rank)
        if rank != dst_rank:
            dest = capi.gather_dest_open(ctx, payload[0])
        dist.barrier()
    if timer:
        timer("push_begin")
    capi.gather_push(ctx, part, dest, row_off[rank], my_byte_off)
    if timer:
        timer("push_end")
    if world > 1:
        dist.barrier()
    if rank == dst_rank:
        capi.gather_dest_finish(ctx, dest)
        return dest
    dest.release()
    return None
