"""N > 1 path on CPU: two processes over gloo (127.0.0.1) run the row-range sharding protocol of SURVEY.md §8(e).

Each rank takes its shard_rows() range of the synthetic table, filters it (here with the CPU oracle standing in for
the per-GPU kernel — the host-side protocol is what is under test), exchanges ONLY its survivor count
(exchange_counts), and rank 0 checks that the rank-ordered concatenation, cut by limit_take(), equals the
single-process result.  No data-path collective is involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rivulus_b200 import capi
from rivulus_b200.sharding import exchange_counts, limit_take, ordered_offsets, shard_rows

N_ROWS = 100_000
SPEC = [("k", capi.SYNTH_KEY1000, 0, 0), ("a", capi.SYNTH_I64, 1, 10), ("b", capi.SYNTH_F64, 2, 0)]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, limit, q):
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = shard_rows(N_ROWS, rank, world)
        df = O.DataFrame.synth(SPEC, e - b, row0=b)
        out = O.LazyFrame.from_dataframe(df).filter(O.col("k").gt(O.lit(499))).collect()
        local = out.to_dict() if out.height() else {"k": [], "a": [], "b": []}
        counts = exchange_counts(out.height())
        take = limit_take(counts, limit)
        offs = ordered_offsets(take)
        part = {k: v[:take[rank]] for k, v in local.items()}
        gathered = [None] * world
        dist.gather_object((rank, offs[rank], part), gathered if rank == 0 else None, dst=0)
        if rank == 0:
            merged = {"k": [], "a": [], "b": []}
            for r, off, p in sorted(gathered):
                assert off == len(merged["k"])              # each shard lands at its exclusive-scan offset
                for k in merged:
                    merged[k].extend(p[k])
            q.put((counts, merged))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("limit", [-1, 1000, 30_000])
def test_two_rank_row_range_sharding(limit):
    from oracle import oracle as O
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, limit, q)) for r in range(world)]
    for p in procs:
        p.start()
    counts, merged = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    df = O.DataFrame.synth(SPEC, N_ROWS, row0=0)
    lf = O.LazyFrame.from_dataframe(df).filter(O.col("k").gt(O.lit(499)))
    want = (lf.limit(limit) if limit >= 0 else lf).collect().to_dict()
    assert sum(counts) == O.LazyFrame.from_dataframe(df).filter(O.col("k").gt(O.lit(499))).collect().height()
    assert merged == want


def test_gather_plan_offsets_match_sequential_concat():
    """plan_gather: row / byte offsets of every shard in the ordered result, with and without a global LIMIT."""
    from rivulus_b200.sharding import plan_gather
    per_rank = [[5, 50, 7], [0, 0, 0], [3, 31, 2], [9, 90, 11]]
    take, row_off, byte_off, total, totals = plan_gather(per_rank)
    assert take == [5, 0, 3, 9] and row_off == [0, 5, 5, 8] and total == 17
    assert byte_off == [[0, 0], [50, 7], [50, 7], [81, 9]] and totals == [171, 20]
    take, row_off, _, total, _ = plan_gather(per_rank, limit=7)
    assert take == [5, 0, 2, 0] and row_off == [0, 5, 5, 7] and total == 7
