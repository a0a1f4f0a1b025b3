"""The full-size property oracle (survivor count + order-sensitive checksums straight from the counter-based generator,
with nulls and strings) is itself checked against the eager oracle engine on small tables (CPU), then used on the GPU at
BASELINE configs[2] / configs[4] shapes (strings + 10 % nulls; Int64/Float64/Boolean with the predicate column projected)."""
import numpy as np
import pytest

from oracle import oracle as O
from rivulus_b200 import capi

M64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15
NULL_TAG = 0x6E756C6C6E756C6C


def splitmix64(x):
    x = (x + GOLDEN) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def term(bits, pos):
    return splitmix64((bits + (pos + 1) * GOLDEN) & M64)


def column_checksum(values):
    s = 0
    for i, v in enumerate(values):
        if v is None: bits = NULL_TAG
        elif isinstance(v, bool): bits = int(v)
        elif isinstance(v, int): bits = v & M64
        elif isinstance(v, float): bits = int(np.float64(v).view(np.uint64))
        else:
            b = v.encode()
            h = len(b)
            for ch in b:
                h = splitmix64(h ^ ch)
            bits = h
        s = (s + term(bits, i)) & M64
    return s


@pytest.mark.parametrize("op,lit", [(">", 700.0), ("<", 150.0), ("!=", 3.0), (">=", None)])
def test_synth_checksum_oracle_matches_eager_engine(op, lit):
    n, row0 = 4000, 12345
    spec = [("x", capi.SYNTH_F64, 0, 10), ("name", capi.SYNTH_STR, 1, 10), ("v", capi.SYNTH_I64, 2, 10), ("f", capi.SYNTH_BOOL, 3, 10)]
    df = O.DataFrame.synth(spec, n, row0=row0)
    meth = {">": "gt", "<": "lt", "!=": "neq", ">=": "gte"}[op]
    out = O.LazyFrame.from_dataframe(df).filter(getattr(O.col("x"), meth)(O.lit(lit))).collect()
    count, sums, nulls, nbytes = O.synth_filter_checksums_nulls(n, row0, (capi.SYNTH_F64, 0, 10), op, lit,
                                                                [(s[1], s[2], s[3]) for s in spec], threads=3)
    assert count == out.height()
    for j, (name, *_rest) in enumerate(spec):
        vals = out.column(name)
        assert sums[j] == column_checksum(vals), name
        assert nulls[j] == sum(v is None for v in vals), name
    assert nbytes[1] == sum(len(v) for v in out.column("name") if v is not None)


def _golden_cases(workload, rows):
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "synth_checksums.json")
    return [c for c in json.load(open(path))["cases"] if c["workload"] == workload and c["rows"] == rows]


def test_golden_checksum_file_matches_the_oracle():
    """tests/golden/synth_checksums.json (scripts/gen_golden_checksums.py) is what the oracle computes: the cheapest full-size case
    (0.1 % of 10^9 rows) is recomputed here, and a prefix property ties the rest: the count of `k > T` over the first 10^8 rows is
    monotone in T and bounded by the committed full-size counts."""
    cases = _golden_cases("configs[1]", 1_000_000_000)
    assert [c["pred"]["literal"] for c in cases] == [998, 899, 499, 99] and len(_golden_cases("configs[4]", 4_000_000_000)) == 2
    c = cases[0]
    count, sums = O.synth_filter_checksums(c["rows"], c["row0"], c["pred"]["kind"], c["pred"]["col_id"], c["pred"]["op"], c["pred"]["literal"],
                                           [tuple(p) for p in c["proj"]])
    assert count == c["count"] and sums == [int(x) for x in c["checksums"]]
    prev = 0
    for c in cases:
        part, _ = O.synth_filter_checksums(100_000_000, 0, 0, 0, ">", c["pred"]["literal"], [(capi.SYNTH_I64, 1)])
        assert prev <= part <= c["count"]
        prev = part


def _gpu_check(ctx, n, row0, pred, op, lit, proj_spec, limit=-1):
    spec = [pred] + list(proj_spec)
    gb = ctx.gen_batch(spec, n, row0)
    got = ctx.filter_project(gb, capi.predicate(0, op, lit), list(range(len(spec))), limit)   # the predicate column is projected too
    count, sums, nulls, nbytes = O.synth_filter_checksums_nulls(n, row0, pred, op, lit, spec, limit=limit)
    assert got.num_rows() == count
    for j in range(len(spec)):
        v = got.view(j)
        assert v.null_count == nulls[j], f"column {j} null count"
        assert (v.validity is not None) == (nulls[j] > 0), f"column {j}: bitmap present iff a survivor is null"
        assert got.checksum(j) == sums[j], f"column {j} checksum"
        if spec[j][0] == capi.SYNTH_STR:
            assert v.data_len == nbytes[j]


@pytest.mark.gpu
@pytest.mark.parametrize("plan", [capi.PLAN_FUSED, capi.PLAN_TWO_PASS])
@pytest.mark.parametrize("op,lit", [(">", 900.0), (">", 500.0), ("<", 100.0)])
def test_config3_strings_and_nulls_at_scale(plan, op, lit):
    """configs[2] shape: filter on Float64, project a variable-length StringArray (avg 24 B) and an Int64, 10 % nulls everywhere."""
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_PLAN, plan)
    try:
        _gpu_check(ctx, 6_000_000, 7_000_000_000, (capi.SYNTH_F64, 0, 10), op, lit, [(capi.SYNTH_STR, 1, 10), (capi.SYNTH_I64, 2, 10)])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_config3_one_full_size_batch():
    """configs[2] at the size of one of its RecordBatches: the LAST 50 M-row batch of the 200 M-row table (1.08 GB of string bytes, the
    largest a batch may hold under int32 offsets is 2 GiB), 10 % nulls in every column, filter on Float64 at 50 %: survivor count,
    order-sensitive checksums, null counts, bitmap rule and surviving string bytes equal the oracle's, under the AUTO plan."""
    ctx = capi.Context(0)
    try:
        _gpu_check(ctx, 50_000_000, 150_000_000, (capi.SYNTH_F64, 0, 10), ">", 500.0, [(capi.SYNTH_STR, 1, 10), (capi.SYNTH_I64, 2, 10)])
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("op,lit,limit", [(">", 899, -1), (">", 499, -1), ("<=", 99, -1), (">", 499, 1_000_000)])
def test_config5_int_float_bool_at_scale(op, lit, limit):
    """configs[4] shape (one shard): {k: Int64, v: Float64, f: Boolean}, every column projected, row range deep inside a 4 B-row table."""
    ctx = capi.Context(0)
    try:
        _gpu_check(ctx, 48_000_000, 3_500_000_000, (capi.SYNTH_KEY1000, 0, 0), op, lit, [(capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)], limit)
        _gpu_check(ctx, 16_000_000, 3_500_000_000, (capi.SYNTH_KEY1000, 0, 5), op, lit, [(capi.SYNTH_F64, 1, 5), (capi.SYNTH_BOOL, 2, 5)], limit)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_config5_one_full_size_shard():
    """configs[4] at the size of one of its shards: the LAST 500 M-row shard (rows 3.5 B .. 4 B) of the 4 B-row {k: Int64, v: Float64,
    f: Boolean} table at 8 GPUs, every column projected, 50 %: count + order-sensitive checksums (the Boolean column included)."""
    ctx = capi.Context(0)
    try:
        _gpu_check(ctx, 500_000_000, 3_500_000_000, (capi.SYNTH_KEY1000, 0, 0), ">", 499, [(capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_full_size_one_billion_rows_properties():
    """BASELINE configs[1] at its full size (10^9 rows x {k, a, b, c, d}): exact count + order-sensitive checksums from the oracle at
    0.1 %, and size-independent properties at 50 %: complement counts add up to N, both execution plans agree bit for bit
    (checksums of every projected column), filtering the result again with the same predicate keeps every row (idempotence),
    and the ordered concatenation of two row-range shards equals the whole."""
    n = 1_000_000_000
    spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_I64, 1, 0), (capi.SYNTH_F64, 2, 0), (capi.SYNTH_I64, 3, 0), (capi.SYNTH_F64, 4, 0)]
    ctx = capi.Context(0)
    try:
        table = ctx.gen_batch(spec, n)
        # exact, against the generator-derived oracle
        out = ctx.filter_project(table, capi.predicate(0, ">", 998), [1, 2, 3, 4])
        count, sums = O.synth_filter_checksums(n, 0, capi.SYNTH_KEY1000, 0, ">", 998, [(s[0], s[1]) for s in spec[1:]])
        assert out.num_rows() == count and [out.checksum(j) for j in range(4)] == sums
        out.release()
        # exact at every selectivity of the sweep, against the values the same oracle computed offline (tests/golden/synth_checksums.json)
        for case in _golden_cases("configs[1]", n):
            o = ctx.filter_project(table, capi.predicate(0, case["pred"]["op"], case["pred"]["literal"]), [1, 2, 3, 4])
            assert o.num_rows() == case["count"], case["pred"]
            assert [o.checksum(j) for j in range(4)] == [int(x) for x in case["checksums"]], case["pred"]
            o.release()
        # properties at 50 %
        res = {}
        for plan in (capi.PLAN_TWO_PASS, capi.PLAN_FUSED):
            ctx.set_option(capi.OPT_PLAN, plan)
            o = ctx.filter_project(table, capi.predicate(0, ">", 499), [0, 1, 2, 3, 4])
            res[plan] = (o.num_rows(), [o.checksum(j) for j in range(5)])
            if plan == capi.PLAN_FUSED:
                again = ctx.filter_project(o, capi.predicate(0, ">", 499), [0, 1, 2, 3, 4])
                assert (again.num_rows(), [again.checksum(j) for j in range(5)]) == res[plan]      # idempotent
                again.release()
            o.release()
        assert res[capi.PLAN_TWO_PASS] == res[capi.PLAN_FUSED]
        ctx.set_option(capi.OPT_PLAN, capi.PLAN_AUTO)
        comp = ctx.filter_project(table, capi.predicate(0, "<=", 499), [0])
        assert comp.num_rows() + res[capi.PLAN_FUSED][0] == n
        comp.release()
        # two row-range shards (rvl_shard_range), concatenated in rank order, equal the whole (checked on the narrow 10 % query)
        whole = ctx.filter_project(table, capi.predicate(0, ">", 899), [2])
        parts = []
        for r in range(2):
            b, e = capi.shard_range(n, r, 2)
            parts.append(ctx.filter_project(table.slice(b, e - b), capi.predicate(0, ">", 899), [2]))
        cat = ctx.concat(parts)
        assert cat.num_rows() == whole.num_rows() and cat.checksum(0) == whole.checksum(0)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_offset_overflow_is_reported_not_wrapped():
    """The one deliberate deviation from the reference: where `StringArray::new` wraps int32 offsets silently (string.rs:31),
    upload / concat / take report RVL_OFFSET_OVERFLOW (status 7)."""
    import ctypes as C
    ctx = capi.Context(0)
    try:
        # upload: a column whose data buffer exceeds the int32 range is refused before anything is copied
        off = np.zeros(2, np.int32)
        col = capi.Column(capi.STRING, 1, 0, None, None, off, np.zeros(1, np.uint8)).as_struct()
        col.data_len = (1 << 31) + 5
        arr = (capi.RvlColumn * 1)(col)
        out = C.c_void_p()
        assert capi.lib().rvl_batch_upload(ctx._h, arr, 1, C.byref(out)) == capi.OFFSET_OVERFLOW
        # concat: two batches of ~1.1 GB of string bytes each are fine alone, their concatenation is not
        spec = [(capi.SYNTH_STR, 1, 0)]
        a = ctx.gen_batch(spec, 46_000_000, 0)
        b = ctx.gen_batch(spec, 46_000_000, 46_000_000)
        assert a.view(0).data_len + b.view(0).data_len > (1 << 31)
        with pytest.raises(capi.RivulusError) as ei:
            ctx.concat([a, b])
        assert ei.value.status == capi.OFFSET_OVERFLOW
        # ... while a concatenation that references fewer bytes than the buffers hold succeeds (data_len is only an upper bound)
        ok = ctx.concat([a.slice(0, 1000), b.slice(5, 1000)])
        assert ok.num_rows() == 2000
        b.release()
        # take: indices that repeat rows until the gathered bytes pass 2 GiB (the wrapped int32 total would be positive again)
        small = a.slice(0, 1_000_000)
        idx = np.tile(np.arange(1_000_000, dtype=np.int64), 190)     # ~190 x 24 MB = 4.5 GB
        with pytest.raises(capi.RivulusError) as ei:
            small.take(idx)
        assert ei.value.status == capi.OFFSET_OVERFLOW
    finally:
        ctx.close()
