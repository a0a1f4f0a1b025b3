"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.
Bit-exact bar: values (raw 64-bit patterns), bitmaps, string offsets/bytes, null counts, bitmap presence.

Reference behaviour under test (files under /root/reference/src): physical_plan/plan.rs:97-173 (eager
Filter/Select/Limit), execution/record_batch.rs:92-342 (slice/take/filter/concat), execution/stream.rs +
physical_plan/streaming.rs (FilterStream/SelectStream/LimitStream), datatypes/series.rs:87-117 (truth table).
"""
import numpy as np
import pytest

from oracle import oracle as O
from rivulus_b200 import capi
from tests.parity import Col, assert_batches_equal, oracle_batch, random_col, run_cmp, run_mask, upload

pytestmark = pytest.mark.gpu

OPS = ["==", "!=", "<", ">", "<=", ">="]


# every parity test runs under both execution plans of the operator (include/rivulus_gpu.h: rvl_plan) and under AUTO
@pytest.fixture(scope="module", params=["fused", "two_pass", "auto"])
def ctx(request):
    c = capi.Context(0)
    c.set_option(capi.OPT_PLAN, {"fused": capi.PLAN_FUSED, "two_pass": capi.PLAN_TWO_PASS, "auto": capi.PLAN_AUTO}[request.param])
    yield c
    c.close()


# ------------------------------------------------------------------ reference known-answer tests, through the GPU
def fixture_cols():
    # execution/record_batch.rs:594-604: id [1,2,3], name ["Alice", None, "Charlie"], active [T,F,T]
    return [Col("i64", 3, np.array([1, 2, 3])), Col("str", 3, None, np.array([1, 0, 1], bool), [b"Alice", b"", b"Charlie"]),
            Col("bool", 3, np.array([True, False, True]))]


def test_golden_record_batch_filter(ctx):  # record_batch.rs:822-879
    cols = fixture_cols()
    for mask, valid, rows in [([1, 0, 1], None, 2), ([1, 1, 1], None, 3), ([0, 0, 0], None, 0), ([1, 0, 0], [1, 0, 1], 1)]:
        m = Col("bool", 3, np.array(mask, bool), None if valid is None else np.array(valid, bool))
        got = run_mask(ctx, cols + [m], 3, [0, 1, 2], tag="rb.filter")
        assert got.num_rows() == rows
    got = run_mask(ctx, cols + [Col("bool", 3, np.array([1, 0, 1], bool))], 3, [0, 1, 2])
    assert got.download_column(0).to_list() == [1, 3] and got.download_column(1).to_list() == ["Alice", "Charlie"]


def eager_cols():
    # physical_plan/plan.rs:295-327: name, age, score
    return [Col("str", 3, None, None, [b"Alice", b"Bob", b"Charlie"]), Col("i64", 3, np.array([25, 30, 35])),
            Col("f64", 3, np.array([85.5, 92.0, 78.5]))]


def test_golden_eager_filters(ctx):
    cols = eager_cols()
    got = run_cmp(ctx, cols, 1, ">", 25, [0, 1, 2], tag="plan.rs:505-525")
    assert got.download_column(1).to_list() == [30, 35] and got.download_column(0).to_list() == ["Bob", "Charlie"]
    got = run_cmp(ctx, cols, 0, "==", "Bob", [0, 1, 2], tag="plan.rs:528-547")
    assert got.download_column(0).to_list() == ["Bob"]
    got = run_cmp(ctx, cols, 2, "<", 90.0, [0, 1, 2], tag="plan.rs:550-569")
    assert got.download_column(0).to_list() == ["Alice", "Charlie"]
    got = run_cmp(ctx, cols, 1, ">", 100, [0, 1, 2], tag="plan.rs:572-589")
    assert got.num_rows() == 0 and got.num_columns() == 3
    got = run_cmp(ctx, cols, 1, ">", 25, [0, 1, 2], limit=1, tag="plan.rs:672-704")
    assert got.download_column(0).to_list() == ["Bob"]
    got = run_cmp(ctx, cols, 1, ">=", 30, [0, 2], tag="plan.rs:707-735")
    assert got.download_column(0).to_list() == ["Bob", "Charlie"] and got.download_column(1).to_list() == [92.0, 78.5]


def test_golden_limits(ctx):  # plan.rs:615-667
    cols = eager_cols()
    gb = upload(ctx, cols)
    for lim, rows in [(2, 2), (10, 3), (0, 0)]:
        got = ctx.filter_project(gb, capi.true_predicate(), [0, 1, 2], lim)
        assert got.num_rows() == rows
        want = oracle_batch(cols)
        want = O.RecordBatch.concat([want.slice(0, min(lim, 3))])
        assert_batches_equal(got, want, f"limit {lim}")


# ------------------------------------------------------------------ truth table x column types x sizes
@pytest.mark.parametrize("n", [0, 1, 63, 64, 65, 2047, 2048, 2049, 5000])
def test_int64_all_ops_sizes(ctx, n):
    rng = np.random.default_rng(n)
    cols = [random_col(rng, "i64", n, 0.0, lo=0, hi=20), random_col(rng, "i64", n), random_col(rng, "f64", n)]
    for op in OPS:
        run_cmp(ctx, cols, 0, op, 10, [1, 2, 0], tag=f"n={n}")


@pytest.mark.parametrize("op", OPS)
def test_nulls_everywhere_int_float(ctx, op):
    rng = np.random.default_rng(11)
    n = 7001
    cols = [random_col(rng, "i64", n, 0.2, lo=0, hi=50), random_col(rng, "f64", n, 0.3, specials=True),
            random_col(rng, "i64", n, 0.15), random_col(rng, "bool", n, 0.25), random_col(rng, "null", n)]
    for lit in (25, None, 25.0, "x", True):           # same type, Null literal, cross-type literals
        run_cmp(ctx, cols, 0, op, lit, [0, 1, 2, 3, 4], tag="pred=i64")
    for lit in (499.5, float("nan"), 0.0, -0.0, None, 7):
        run_cmp(ctx, cols, 1, op, lit, [1, 0, 3], tag="pred=f64")


@pytest.mark.parametrize("op", OPS)
def test_boolean_string_null_predicate_columns(ctx, op):
    rng = np.random.default_rng(5)
    n = 4500
    cols = [random_col(rng, "bool", n, 0.2), random_col(rng, "str", n, 0.2, maxlen=4, alphabet=b"ab"), random_col(rng, "null", n),
            random_col(rng, "i64", n, 0.1), random_col(rng, "str", n, 0.3, maxlen=30, alphabet=b"xyz0123456789")]
    for lit in (True, False, None, 1):
        run_cmp(ctx, cols, 0, op, lit, [0, 3, 4], tag="pred=bool")
    for lit in ("ab", "", "b", "abab", None, 3.5):
        run_cmp(ctx, cols, 1, op, lit, [1, 3, 0], tag="pred=str")
    for lit in (None, 1, "a"):
        run_cmp(ctx, cols, 2, op, lit, [2, 3], tag="pred=nullcol")


def test_all_null_survivors_and_no_nulls_bitmap_rule(ctx):
    # bitmap present iff a SURVIVOR is null (primitive.rs:180-185); placeholder 0 under nulls (record_batch.rs:144)
    n = 300
    k = np.arange(n)
    v = np.arange(n) * 10
    valid = np.ones(n, bool); valid[200:] = False
    cols = [Col("i64", n, k), Col("i64", n, v, valid), Col("f64", n, v * 0.5, valid), Col("bool", n, (k % 2 == 0), valid)]
    got = run_cmp(ctx, cols, 0, "<", 100, [1, 2, 3], tag="no null survivor")     # survivors all valid -> no bitmaps
    assert all(got.view(j).validity is None for j in range(3))
    got = run_cmp(ctx, cols, 0, ">=", 250, [1, 2, 3], tag="all null survivors")
    c = got.download_column(0)
    assert c.null_count == 50 and not c.values.any()


def test_sliced_views_with_bit_offsets(ctx):
    # slices are (buffer, offset, length) views with arbitrary bit offsets (bitmap.rs:104-112, primitive.rs:107-117)
    rng = np.random.default_rng(3)
    for off in (1, 7, 8, 31, 33, 64, 67, 129):
        n = 3000
        cols = [random_col(rng, "i64", n, 0.2, offset=off, tail=5, lo=0, hi=100), random_col(rng, "f64", n, 0.2, offset=off + 3, tail=9),
                random_col(rng, "bool", n, 0.2, offset=off + 1, tail=2), random_col(rng, "str", n, 0.2, offset=off, tail=1)]
        run_cmp(ctx, cols, 0, ">", 50, [0, 1, 2, 3], tag=f"off={off}")
        run_cmp(ctx, cols, 1, "<=", 500.0, [3, 2, 1, 0], limit=777, tag=f"off={off}")
        run_mask(ctx, cols, 2, [0, 1, 2, 3], tag=f"off={off}")


def test_device_side_slice_then_filter(ctx):
    rng = np.random.default_rng(8)
    n = 10000
    cols = [random_col(rng, "i64", n, 0.1, lo=0, hi=100), random_col(rng, "f64", n, 0.1), random_col(rng, "bool", n, 0.1),
            random_col(rng, "str", n, 0.1)]
    gb, ob = upload(ctx, cols), oracle_batch(cols)
    for off, ln in [(0, n), (1, 100), (37, 5000), (4099, 2048), (n - 1, 1), (500, 0)]:
        got = ctx.filter_project(gb.slice(off, ln), capi.predicate(0, ">=", 40), [3, 2, 1, 0])
        want = ob.slice(off, ln).filter_project_cmp(0, ">=", 40, [3, 2, 1, 0])
        assert_batches_equal(got, want, f"slice({off},{ln})")
    with pytest.raises(capi.RivulusError) as ei:
        gb.slice(n - 1, 2)
    assert ei.value.status == capi.OUT_OF_BOUNDS and "Slice out of bounds" in ei.value.message


@pytest.mark.parametrize("sel", [0.001, 0.1, 0.5, 0.9, 1.0])
def test_million_rows_selectivity(ctx, sel):
    rng = np.random.default_rng(int(sel * 1000))
    n = 1_000_003
    cols = [Col("i64", n, rng.integers(0, 1000, n)), random_col(rng, "i64", n, 0.1, lo=-2**62, hi=2**62), random_col(rng, "f64", n),
            random_col(rng, "i64", n, lo=-2**62, hi=2**62), random_col(rng, "f64", n, 0.05), random_col(rng, "bool", n, 0.1)]
    thr = int(round(1000 * (1 - sel))) - 1
    run_cmp(ctx, cols, 0, ">", thr, [1, 2, 3, 4, 5], tag=f"sel={sel}")


def test_many_columns_multi_launch(ctx):
    # > 8 fixed-width and > 16 bit-packed columns: replays the selection bitmap over several launches
    rng = np.random.default_rng(21)
    n = 6000
    cols = [random_col(rng, "i64", n, lo=0, hi=10)]
    for i in range(11):
        cols.append(random_col(rng, "i64" if i % 2 else "f64", n, 0.2))
    for i in range(9):
        cols.append(random_col(rng, "bool", n, 0.2))
    proj = list(range(1, len(cols))) + [0]
    run_cmp(ctx, cols, 0, ">", 4, proj, tag="21 cols")
    run_cmp(ctx, cols, 0, ">", 4, proj, limit=1234, tag="21 cols limit")


@pytest.mark.parametrize("limit", [0, 1, 31, 32, 33, 1000, 2048, 2049, 100000])
def test_limit_semantics(ctx, limit):
    rng = np.random.default_rng(limit)
    n = 50_000
    cols = [random_col(rng, "i64", n, 0.1, lo=0, hi=100), random_col(rng, "f64", n, 0.1), random_col(rng, "bool", n, 0.1),
            random_col(rng, "str", n, 0.1, maxlen=20)]
    run_cmp(ctx, cols, 0, ">", 49, [0, 1, 2, 3], limit=limit, tag="limit")
    run_mask(ctx, cols, 2, [3, 1], limit=limit, tag="limit mask")


def test_strings_long_and_empty(ctx):
    rng = np.random.default_rng(99)
    n = 9000
    strings = []
    for i in range(n):
        L = [0, 1, 5, 24, 40, 200, 1500][int(rng.integers(0, 7))]
        strings.append(bytes(rng.integers(97, 123, L).astype(np.uint8)))
    valid = rng.random(n) > 0.1
    cols = [random_col(rng, "f64", n, 0.1), Col("str", n, None, valid, strings), Col("str", n, None, None, ["ü€🦀".encode()] * n)]
    run_cmp(ctx, cols, 0, ">", 300.0, [1, 2, 0], tag="strings")
    run_cmp(ctx, cols, 0, "<", 300.0, [1], tag="strings nulls pass <")
    # medium strings (0..90 bytes, every alignment phase): the one-lane-per-string word copy, tiles spanning several staging chunks
    mid = [bytes(rng.integers(97, 123, int(rng.integers(0, 91))).astype(np.uint8)) for _ in range(n)]
    cols = [random_col(rng, "f64", n, 0.1), Col("str", n, None, valid, mid)]
    for op, lit in ((">", 100.0), (">", 700.0), ("<", 980.0)):
        run_cmp(ctx, cols, 0, op, lit, [1], tag=f"medium strings {op} {lit}")


def test_predicate_mask_kernel(ctx):
    rng = np.random.default_rng(4)
    n = 70_001
    cols = [random_col(rng, "i64", n, 0.2, lo=0, hi=10), random_col(rng, "f64", n, 0.2)]
    gb = upload(ctx, cols)
    for op in OPS:
        m = ctx.predicate_mask(gb, capi.predicate(0, op, 5)).download_column(0)
        bits = capi.unpack_bits(m.values, n)
        want = [O.eval_cmp(None if not cols[0].valid[i] else int(cols[0].values[i]), op, 5) for i in range(0, n, 997)]
        assert [bool(b) for b in bits[::997]] == want


# ------------------------------------------------------------------ batch ops: concat / select / download
def test_concat_matches_reference(ctx):  # record_batch.rs:245-342, tests :882-949
    rng = np.random.default_rng(17)
    parts = []
    for n, nf in [(100, 0.0), (3000, 0.3), (0, 0.0), (65, 0.5), (2048, 0.0)]:
        parts.append([random_col(rng, "i64", n, nf, offset=3, tail=2), random_col(rng, "f64", n, nf), random_col(rng, "bool", n, nf, offset=5),
                      random_col(rng, "str", n, nf, offset=1), random_col(rng, "null", n)])
    got = ctx.concat([upload(ctx, p) for p in parts])
    want = O.RecordBatch.concat([oracle_batch(p) for p in parts])
    assert_batches_equal(got, want, "concat")
    novalid = [[random_col(rng, "i64", 10)], [random_col(rng, "i64", 20)]]
    got = ctx.concat([upload(ctx, p) for p in novalid])
    assert got.view(0).validity is None
    assert_batches_equal(got, O.RecordBatch.concat([oracle_batch(p) for p in novalid]), "concat no nulls")
    with pytest.raises(capi.RivulusError) as ei:
        ctx.concat([upload(ctx, [random_col(rng, "i64", 4)]), upload(ctx, [random_col(rng, "f64", 4)])])
    assert ei.value.status == capi.SCHEMA_MISMATCH and "All batches must have the same schema" in ei.value.message
    with pytest.raises(capi.RivulusError) as ei:
        ctx.concat([])
    assert "Cannot concatenate empty batch list" in ei.value.message


def test_select_and_errors(ctx):
    cols = fixture_cols()
    gb = upload(ctx, cols)
    s = gb.select([2, 0])
    assert s.num_columns() == 2 and s.download_column(1).to_list() == [1, 2, 3]
    with pytest.raises(capi.RivulusError) as ei:
        gb.select([0, 5])
    assert ei.value.status == capi.OUT_OF_BOUNDS and "Column index 5 out of bounds for 3 columns" in ei.value.message
    with pytest.raises(capi.RivulusError) as ei:
        ctx.filter_project(gb, capi.mask_predicate(0), [0])          # stream.rs:147-153
    assert ei.value.status == capi.TYPE_MISMATCH
    with pytest.raises(capi.RivulusError) as ei:
        ctx.upload([cols[0].gpu(), Col("i64", 2, np.array([1, 2])).gpu()])   # record_batch.rs:33-38
    assert ei.value.status == capi.LENGTH_MISMATCH and "Column 1 has length 2 but expected 3" in ei.value.message
    p = capi.predicate(0, "==", 1); p.op = capi.OPS["+"]
    with pytest.raises(capi.RivulusError) as ei:
        ctx.filter_project(gb, p, [0])                                # plan.rs:121-127
    assert ei.value.status == capi.INVALID_OPERATION


# ------------------------------------------------------------------ streaming executor vs the reference's stream chain
@pytest.mark.parametrize("limit", [-1, 0, 1, 1000, 5000, 10**7])
@pytest.mark.parametrize("pinned", [False, True])
def test_stream_filter_select_limit(ctx, limit, pinned):
    rng = np.random.default_rng(limit + 7)
    batches = []
    for n in [4096, 1, 5000, 0, 2048, 3333]:
        batches.append([random_col(rng, "i64", n, 0.1, lo=0, hi=100), random_col(rng, "f64", n, 0.1), random_col(rng, "bool", n, 0.2),
                        random_col(rng, "str", n, 0.1, maxlen=16)])
    dtypes = [capi.INT64, capi.FLOAT64, capi.BOOLEAN, capi.STRING]
    st = ctx.open_stream(dtypes, capi.mask_predicate(2), [0, 3, 1], limit, batch_rows=8192, n_staging=2)
    keep = []
    for b in batches:
        gcols = [c.gpu() for c in b]
        if pinned:   # page-locked sources are copied straight from the caller's buffers (no staging memcpy)
            for gc in gcols:
                for f in ("values", "validity", "offsets", "data"):
                    a = getattr(gc, f)
                    if a is not None and a.size:
                        v, owner = capi.pinned_like(a)
                        setattr(gc, f, v)
                        keep.append(owner)
        st.push(gcols)
    got = st.collect()
    st.close()
    plan = O.StreamingPhysicalPlan.memory_source([oracle_batch(b) for b in batches]).filter("c2").select(["c0", "c3", "c1"])
    if limit >= 0:
        plan = plan.limit(limit)
    assert_batches_equal(got, plan.collect(), f"stream limit={limit}")


def test_stream_limit_stops_transfers(ctx):
    rng = np.random.default_rng(2)
    n = 65536
    st = ctx.open_stream([capi.INT64, capi.INT64], capi.predicate(0, ">", 899), [1], 1000, batch_rows=n, n_staging=2)
    accepted = 0
    cols_all = []
    for i in range(16):
        cols = [Col("i64", n, rng.integers(0, 1000, n)), random_col(rng, "i64", n)]
        cols_all.append(cols)
        accepted += bool(st.push([c.gpu() for c in cols]))
    got = st.collect()
    stats = st.stats()
    st.close()
    assert got.num_rows() == 1000
    assert stats["batches_pushed"] == accepted <= 3 and stats["batches_skipped"] >= 13   # ideal 1; pipeline depth 2 allows <= 3
    ob = O.RecordBatch.concat([oracle_batch(c) for c in cols_all[:2]]).filter_project_cmp(0, ">", 899, [1], 1000)
    assert_batches_equal(got, ob, "stream early stop")


@pytest.mark.parametrize("plan", [capi.PLAN_FUSED, capi.PLAN_TWO_PASS])
@pytest.mark.parametrize("limit", [-1, 700])
@pytest.mark.parametrize("op,lit", [(">", 950), (">", 500), ("<=", 100)])
def test_stream_zero_copy_transfer(plan, limit, op, lit):
    """rvl_transfer = ZERO_COPY: projected fixed-width columns are read in place from pinned host memory (the predicate column and
    strings are staged, the unused column never crosses the bus); same rows, order, nulls and bitmap rule as the staged path and
    as the oracle.  Sliced pushes (bit offsets that are not multiples of 8) included."""
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_PLAN, plan)
    try:
        rng = np.random.default_rng(1234 + limit)
        batches, keep = [], []
        for n, off in [(40_000, 0), (1, 5), (65_536, 0), (0, 0), (12_345, 77), (30_000, 64 * 3 + 13)]:
            # off > 0: the pushed columns are (offset, length) windows of larger pinned buffers, bit offsets not multiples of 8
            batches.append([random_col(rng, "i64", n, 0.1, offset=off, tail=9, lo=0, hi=1000), random_col(rng, "f64", n, 0.1, offset=off, tail=9),
                            random_col(rng, "bool", n, 0.2, offset=off, tail=9), random_col(rng, "i64", n, 0.0, offset=off, tail=9),
                            random_col(rng, "str", n, 0.1, offset=off, tail=9, maxlen=12), random_col(rng, "f64", n, 0.5, offset=off, tail=9)])
        dtypes = [capi.INT64, capi.FLOAT64, capi.BOOLEAN, capi.INT64, capi.STRING, capi.FLOAT64]
        proj = [1, 2, 3, 4, 0]        # column 5 is never looked at; column 0 is predicate + projected (staged once)
        outs = {}
        for mode in (capi.TRANSFER_STAGED, capi.TRANSFER_ZERO_COPY):
            st = ctx.open_stream(dtypes, capi.predicate(0, op, lit), proj, limit, batch_rows=65_536, n_staging=2, transfer=mode)
            for b in batches:
                gcols = [c.gpu() for c in b]
                for gc in gcols:
                    for f in ("values", "validity", "offsets", "data"):
                        a = getattr(gc, f)
                        if a is not None and a.size:
                            v, owner = capi.pinned_like(a)
                            setattr(gc, f, v)
                            keep.append(owner)
                st.push(gcols)
            outs[mode] = st.collect()
            stats = st.stats()
            st.close()
            if mode == capi.TRANSFER_ZERO_COPY and limit < 0:
                staged_rows = sum(b[0].length for b in batches)
                assert stats["h2d_bytes"] < staged_rows * (8 + 4 + 12 + 2), "only the predicate column and the strings are staged"
        want = O.RecordBatch.concat([oracle_batch(b) for b in batches if b[0].length > 0]).filter_project_cmp(0, op, lit, proj, limit)
        assert_batches_equal(outs[capi.TRANSFER_ZERO_COPY], want, f"zero-copy stream {op} {lit} limit={limit}")
        assert_batches_equal(outs[capi.TRANSFER_STAGED], want, f"staged stream {op} {lit} limit={limit}")
    finally:
        ctx.close()


@pytest.mark.parametrize("limit", [-1, 30_000])
def test_stream_many_batches_pushed_before_draining(ctx, limit):
    """every in-flight launch owns its pinned mailbox slot, so pushing far more batches than any fixed ring
    held (200 here, one launch each) before the first drain returns the same rows as the oracle — no aliased counters.  Beyond
    32 outstanding launches push() retires the oldest ones itself and merges their outputs 16 at a time."""
    rng = np.random.default_rng(99)
    n_batches, n = 200, 1024
    batches = [[Col("i64", n, rng.integers(0, 1000, n)), random_col(rng, "f64", n, 0.1), random_col(rng, "bool", n, 0.1)] for _ in range(n_batches)]
    st = ctx.open_stream([capi.INT64, capi.FLOAT64, capi.BOOLEAN], capi.predicate(0, ">", 799), [1, 0, 2], limit, batch_rows=n, n_staging=3)
    for b in batches:
        st.push([c.gpu() for c in b])
    assert limit >= 0 or st.launches() == n_batches
    got = st.collect()
    st.close()
    want = O.RecordBatch.concat([oracle_batch(b) for b in batches]).filter_project_cmp(0, ">", 799, [1, 0, 2], limit)
    assert_batches_equal(got, want, f"200 pushed batches limit={limit}")


@pytest.mark.parametrize("transfer", [capi.TRANSFER_STAGED, capi.TRANSFER_ZERO_COPY, capi.TRANSFER_AUTO])
@pytest.mark.parametrize("limit", [-1, 5_000])
def test_stream_coalesces_small_batches(transfer, limit):
    """Batches smaller than the slot are appended to the open group: one operator launch per group, same rows / order / nulls as
    one launch per batch and as the oracle.  Windows of two big pinned tables (adjacent, so in-place columns can join a group),
    a ragged batch that forces a flush, validity that appears mid-stream, an empty batch, and next() interleaved with push()."""
    ctx = capi.Context(0)
    try:
        rng = np.random.default_rng(4242 + limit)
        rows = 40 * 4096 + 777
        k = rng.integers(0, 1000, rows).astype(np.int64)
        a = rng.integers(-2**62, 2**62, rows).astype(np.int64)
        b = rng.random(rows) * 1000.0
        f = rng.random(rows) < 0.5
        va = rng.random(rows) > 0.1
        va[:12 * 4096] = True                   # the first windows of `a` carry no validity buffer at all
        keep = []

        def pin(x):
            v, owner = capi.pinned_like(x)
            keep.append(owner)
            return v
        pk, pa, pb = pin(k), pin(a), pin(b)
        pf, pva = pin(capi.pack_bits(f)), pin(capi.pack_bits(va))
        sizes = [4096] * 12 + [4096] * 10 + [1000, 0, 4096, 4096, 64, 4096 * 3, 4096 * 8] + [4096] * 3
        sizes.append(rows - sum(sizes) - 13)     # a window that does not start on a 64-row boundary comes last
        windows, off = [], 0
        for i, n in enumerate(sizes):
            if i == len(sizes) - 1:
                off += 13
            windows.append((off, n))
            off += n
        assert off == rows
        dtypes = [capi.INT64, capi.INT64, capi.FLOAT64, capi.BOOLEAN]

        def cols_of(o, n):
            with_valid = o >= 12 * 4096
            return [capi.Column(capi.INT64, n, o, pk), capi.Column(capi.INT64, n, o, pa, pva if with_valid else None),
                    capi.Column(capi.FLOAT64, n, o, pb), capi.Column(capi.BOOLEAN, n, o, pf)]

        outs = {}
        for cap in (4096 * 8, 4096 * 3):         # capacity of a slot: groups of up to 8 / 3 batches
            st = ctx.open_stream(dtypes, capi.predicate(0, ">", 699), [1, 2, 3, 0], limit, batch_rows=max(cap, max(sizes)), n_staging=3, transfer=transfer)
            got_parts = []
            for i, (o, n) in enumerate(windows):
                st.push(cols_of(o, n))
                if i == 20:                      # a consumer that drains mid-stream: flushes the open group if nothing else is pending
                    while True:
                        bt = st.next_batch()
                        if bt is None:
                            break
                        got_parts.append(bt)
            launches = st.launches()
            got_parts.append(st.collect())
            st.close()
            assert limit >= 0 or launches < len(sizes) // 2, f"batches were not coalesced: {launches} launches for {len(sizes)} pushes"
            outs[cap] = ctx.concat(got_parts) if len(got_parts) > 1 else got_parts[0]
        whole = [Col("i64", rows, k), Col("i64", rows, a, va), Col("f64", rows, b), Col("bool", rows, f)]
        parts = []
        for (o, n) in windows:
            if n > 0:
                parts.append(oracle_batch([Col(c.dtype, n, c.values[o:o + n], None if c.valid is None else c.valid[o:o + n]) for c in whole]))
        want = O.RecordBatch.concat(parts).filter_project_cmp(0, ">", 699, [1, 2, 3, 0], limit)
        for cap, got in outs.items():
            assert_batches_equal(got, want, f"coalesced stream cap={cap} transfer={transfer} limit={limit}")
    finally:
        ctx.close()


@pytest.mark.parametrize("exact,overlap", [(0, 0), (1, 0), (0, 1), (1, 1)])
def test_two_pass_exact_allocation_and_forked_bits(exact, overlap):
    """RVL_OPT_EXACT_ALLOC (outputs sized from the scan's count / the string sizes pass) and RVL_OPT_BITS_OVERLAP (bit-packed
    compaction on the forked stream) change where and when the bytes are written, never which bytes."""
    rng = np.random.default_rng(17)
    n = 300_123
    cols = [random_col(rng, "i64", n, 0.1, lo=0, hi=1000), random_col(rng, "f64", n, 0.2), random_col(rng, "bool", n, 0.1, offset=9),
            random_col(rng, "str", n, 0.1, maxlen=18), random_col(rng, "i64", n, 0.0)]
    c = capi.Context(0)
    c.set_option(capi.OPT_PLAN, capi.PLAN_TWO_PASS)
    c.set_option(capi.OPT_EXACT_ALLOC, exact)
    c.set_option(capi.OPT_BITS_OVERLAP, overlap)
    try:
        for op, lit, limit in ((">", 899, -1), (">", 499, -1), ("<", 3, -1), (">", 2000, -1), (">", 499, 1234)):
            run_cmp(c, cols, 0, op, lit, [3, 1, 2, 4, 0], limit, tag=f"exact={exact} overlap={overlap} {op}{lit} limit={limit}")
    finally:
        c.close()


# ------------------------------------------------------------------ sharded (row-range) execution on one GPU
def test_sharded_two_contexts_one_gpu(ctx):
    rng = np.random.default_rng(31)
    n = 200_000
    cols = [random_col(rng, "i64", n, 0.05, lo=0, hi=1000), random_col(rng, "f64", n, 0.05), random_col(rng, "bool", n, 0.05)]
    ob = oracle_batch(cols)
    ctx2 = capi.Context(0)
    whole = upload(ctx, cols)
    for limit in (-1, 5000, 150_000):
        shards, ranges = [], []
        for r in range(2):
            b, e = capi.shard_range(n, r, 2)
            assert b % 64 == 0
            ranges.append((b, e))
            shards.append(whole.slice(b, e - b))
        outs, counts = capi.filter_project_sharded([ctx, ctx2], shards, capi.predicate(0, ">", 499), [0, 1, 2], limit)
        got = ctx.concat(outs)
        want = ob.filter_project_cmp(0, ">", 499, [0, 1, 2], limit)
        assert sum(counts) == want.num_rows()
        assert_batches_equal(got, want, f"sharded limit={limit}")
        assert_batches_equal(capi.gather_to(ctx2, outs), want, f"sharded + gather_to limit={limit}")
    ctx2.close()


def test_sharded_across_gpus_with_peer_gather():
    """configs[4] for real: every visible GPU owns a contiguous row range ({k: Int64, v: Float64, f: Boolean, s: String} with nulls),
    one process drives one context per GPU, and the ordered result is gathered onto GPU 0 by concat kernels that read the peers'
    buffers over NVLink.  Needs >= 2 GPUs (gpurun --gpus 2); the same code path runs on one GPU in the test above."""
    g = capi.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    g = min(g, 8)
    rng = np.random.default_rng(77)
    n = 300_007
    cols = [random_col(rng, "i64", n, 0.05, lo=0, hi=1000), random_col(rng, "f64", n, 0.05), random_col(rng, "bool", n, 0.05),
            random_col(rng, "str", n, 0.1, maxlen=20)]
    ob = oracle_batch(cols)
    ctxs = [capi.Context(d) for d in range(g)]
    try:
        for limit in (-1, 40_000):
            shards = []
            for r in range(g):
                b, e = capi.shard_range(n, r, g)
                part = [Col(c.dtype, e - b, c.values[b:e] if c.values is not None else None, c.valid[b:e] if c.valid is not None else None,
                            c.strings[b:e] if c.strings is not None else None) for c in cols]
                shards.append(upload(ctxs[r], part))
            outs, counts = capi.filter_project_sharded(ctxs, shards, capi.predicate(0, ">", 499), [3, 0, 1, 2], limit)
            want = ob.filter_project_cmp(0, ">", 499, [3, 0, 1, 2], limit)
            assert sum(counts) == want.num_rows()
            assert_batches_equal(capi.gather_to(ctxs[0], outs), want, f"{g}-GPU sharded + peer gather limit={limit}")
    finally:
        for c in ctxs:
            c.close()


# ------------------------------------------------------------------ full-size properties (count + order-sensitive checksums)
@pytest.mark.parametrize("thr,n", [(998, 64_000_000), (499, 64_000_000), (99, 32_000_000)])
def test_large_synthetic_checksums(ctx, thr, n):
    spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_I64, 1, 0), (capi.SYNTH_F64, 2, 0), (capi.SYNTH_I64, 3, 0), (capi.SYNTH_F64, 4, 0),
            (capi.SYNTH_BOOL, 5, 0)]
    gb = ctx.gen_batch(spec, n)
    got = ctx.filter_project(gb, capi.predicate(0, ">", thr), [1, 2, 3, 4, 5])
    count, sums = O.synth_filter_checksums(n, 0, capi.SYNTH_KEY1000, 0, ">", thr, [(s[0], s[1]) for s in spec[1:]])
    assert got.num_rows() == count
    assert [got.checksum(j) for j in range(5)] == sums


# ------------------------------------------------------------------ two-pass plan: dense / sparse tile classification
@pytest.mark.parametrize("sparse_max", [0, 1, 7, 96, 128, 255, 256])
@pytest.mark.parametrize("limit", [-1, 12345])
def test_two_pass_mixed_density_tiles(sparse_max, limit):
    """Clustered survivors: empty, sparse and dense 2048-row tiles in one batch, nulls in every column, ragged tail,
    bit-offset views; every classification threshold must give the same bytes as the oracle."""
    rng = np.random.default_rng(77 + sparse_max)
    n = 70_001
    k = rng.integers(0, 1000, n)
    dens = np.repeat(rng.choice([0.0, 0.002, 0.03, 0.3, 0.95], size=(n + 2047) // 2048), 2048)[:n]
    k = np.where(rng.random(n) < dens, 2000, k % 500).astype(np.int64)
    cols = [Col("i64", n, k, rng.random(n) > 0.05), random_col(rng, "f64", n, 0.1, offset=3, tail=5), random_col(rng, "i64", n, 0.0),
            random_col(rng, "bool", n, 0.2, offset=13), random_col(rng, "f64", n, 0.5)]
    c = capi.Context(0)
    c.set_option(capi.OPT_PLAN, capi.PLAN_TWO_PASS)
    c.set_option(capi.OPT_SPARSE_MAX, sparse_max)
    try:
        c.set_option(capi.OPT_DENSE_WARPS, 16 if sparse_max % 2 else 8)
        run_cmp(c, cols, 0, ">", 1000, [1, 2, 3, 4, 0], limit, tag=f"two-pass sparse_max={sparse_max} limit={limit}")
        run_cmp(c, cols, 0, "<", 100, [4, 3], limit, tag=f"two-pass nulls-pass sparse_max={sparse_max}")
    finally:
        c.close()


@pytest.mark.parametrize("slots,per_sm,scan_warps,scan_slots,dense_warps",
                         [(2, 1, 8, 1, 8), (3, 2, 16, 1, 8), (6, 2, 16, 3, 8), (14, 1, 8, 2, 16), (12, 1, 16, 2, 16), (2, 1, 16, 2, 16), (14, 1, 32, 3, 16), (8, 1, 32, 1, 16)])
def test_two_pass_ring_depths(slots, per_sm, scan_warps, scan_slots, dense_warps):
    rng = np.random.default_rng(5)
    n = 1_500_000
    cols = [random_col(rng, "i64", n, 0.0, lo=0, hi=1000)] + [random_col(rng, "f64" if j % 2 else "i64", n, 0.1 if j == 1 else 0.0) for j in range(5)]
    c = capi.Context(0)
    c.set_option(capi.OPT_PLAN, capi.PLAN_TWO_PASS)
    c.set_option(capi.OPT_DENSE_SLOTS, slots)
    c.set_option(capi.OPT_DENSE_CTAS_PER_SM, per_sm)
    c.set_option(capi.OPT_SCAN_WARPS, scan_warps)
    c.set_option(capi.OPT_DENSE_WARPS, dense_warps)
    c.set_option(capi.OPT_SCAN_SLOTS, scan_slots)
    try:
        run_cmp(c, cols, 0, ">", 299, [1, 2, 3, 4, 5], tag=f"ring slots={slots} per_sm={per_sm} scan={scan_warps}x{scan_slots}")
        # nulls in the predicate column, Float64 predicate, bitmap predicate, unaligned view, LIMIT: every scan path
        fcols = [random_col(rng, "f64", 300_001, 0.1, offset=1, specials=True), random_col(rng, "bool", 300_001, 0.1, offset=5),
                 random_col(rng, "i64", 300_001, 0.2)]
        run_cmp(c, fcols, 0, "<=", 499.5, [2, 1, 0], tag="scan f64+nulls")
        run_cmp(c, fcols, 1, "==", True, [0, 2], 70_000, tag="scan bits+limit")
        run_mask(c, fcols, 1, [2, 0], tag="scan mask")
    finally:
        c.close()


# ------------------------------------------------------------------ single-pass chunk plan: predicate column projected
@pytest.mark.parametrize("n", [1, 4095, 4096, 4097, 12_289, 300_001, 1_300_000])
@pytest.mark.parametrize("chunk", [2, 1, 0])
def test_chunk_plan_predicate_column_projected(n, chunk):
    """`filter(k <op> T)` projecting k itself (every column, as a Filter without Select does): the chunk kernel keeps the predicate
    values in shared memory between the predicate and the compaction (one HBM read) and orders the output by decoupled look-back.
    Same bytes as the oracle with the plan on and off: Int64 / Float64 predicates with nulls and NaN, every selectivity incl. all /
    none, sparse and dense tiles in one batch, 8 numeric columns (ring back-pressure), Boolean + String columns riding along,
    ragged sizes around the 4096-row chunk."""
    rng = np.random.default_rng(n + chunk)
    k = rng.integers(0, 1000, n)
    dens = np.repeat(rng.choice([0.0, 0.01, 0.2, 0.6, 1.0], size=(n + 2047) // 2048), 2048)[:n]
    k = np.where(rng.random(n) < dens, 2000 + k, k % 900).astype(np.int64)
    cols = [Col("i64", n, k, rng.random(n) > 0.05), random_col(rng, "f64", n, 0.1, specials=True), random_col(rng, "i64", n, 0.0),
            random_col(rng, "bool", n, 0.2), random_col(rng, "str", n, 0.1, maxlen=14), random_col(rng, "f64", n, 0.3),
            random_col(rng, "i64", n, 0.0), random_col(rng, "i64", n, 0.1), random_col(rng, "f64", n, 0.0), random_col(rng, "i64", n, 0.0)]
    c = capi.Context(0)
    c.set_option(capi.OPT_PLAN, capi.PLAN_TWO_PASS)
    c.set_option(capi.OPT_CHUNK_PLAN, chunk)
    try:
        allc = list(range(len(cols)))
        run_cmp(c, cols, 0, ">", 1000, allc, tag=f"chunk={chunk} clustered")           # dense and sparse tiles
        run_cmp(c, cols, 0, "<", 450, [0, 2, 3], tag=f"chunk={chunk} nulls pass")       # ~50 % + every null row
        run_cmp(c, cols, 0, ">=", 0, [2, 0, 0, 4], tag=f"chunk={chunk} all valid rows, k twice")
        run_cmp(c, cols, 0, "==", 123456, [0, 1], tag=f"chunk={chunk} none")
        run_cmp(c, cols, 1, "<=", 499.5, [1, 0, 5, 8], tag=f"chunk={chunk} f64 predicate")
        run_cmp(c, cols, 1, "!=", float("nan"), allc, tag=f"chunk={chunk} f64 != NaN")
    finally:
        c.close()


def test_chunk_plan_streamed_and_unaligned_views():
    """The chunk plan inside a chained (streaming) query — the running row count comes in through base_in — and views whose
    predicate values are not 16-byte aligned (the plan steps aside for the two-pass kernels)."""
    rng = np.random.default_rng(8)
    c = capi.Context(0)
    c.set_option(capi.OPT_PLAN, capi.PLAN_TWO_PASS)
    c.set_option(capi.OPT_CHUNK_PLAN, 2)
    try:
        batches = []
        for n in (20_000, 4096, 33_333):
            batches.append([random_col(rng, "i64", n, 0.1, lo=0, hi=1000), random_col(rng, "f64", n, 0.1), random_col(rng, "bool", n, 0.2)])
        st = c.open_stream([capi.INT64, capi.FLOAT64, capi.BOOLEAN], capi.predicate(0, ">", 300), [0, 1, 2], -1, batch_rows=40_000, n_staging=2)
        for b in batches:
            st.push([x.gpu() for x in b])
        got = st.collect()
        st.close()
        want = O.RecordBatch.concat([oracle_batch(b) for b in batches]).filter_project_cmp(0, ">", 300, [0, 1, 2], -1)
        assert_batches_equal(got, want, "chunk plan, chained batches")
        n = 50_001
        cols = [random_col(rng, "i64", n, 0.1, offset=3, tail=2, lo=0, hi=1000), random_col(rng, "f64", n, 0.0, offset=1)]
        run_cmp(c, cols, 0, ">", 500, [0, 1], tag="unaligned predicate view")
    finally:
        c.close()


# ------------------------------------------------------------------ RecordBatch::take (record_batch.rs:108-178)
def test_take_matches_reference(ctx):
    rng = np.random.default_rng(9)
    n = 10_000
    cols = [random_col(rng, "i64", n, 0.1, offset=3), random_col(rng, "f64", n, 0.0), random_col(rng, "bool", n, 0.2, offset=7),
            random_col(rng, "str", n, 0.15, maxlen=20), Col("null", n)]
    gb, ob = upload(ctx, cols), oracle_batch(cols)
    for idx in ([], [0], [n - 1, 0, n - 1], rng.integers(0, n, 5000).tolist(), list(range(n - 1, -1, -1))):
        assert_batches_equal(gb.take(idx), ob.take(idx), f"take {len(idx)} rows")
    with pytest.raises(capi.RivulusError, match=f"Index {n} out of bounds for {n} rows"):
        gb.take([0, n, 1])


# ------------------------------------------------------------------ BooleanArray::{and, or, not, count_true} (array/boolean.rs:120-178)
def _bool_batch(ctx, vals):
    return ctx.upload([capi.Column.from_list(vals, capi.BOOLEAN)])


def test_golden_boolean_logical_ops(ctx):  # boolean.rs:626-691
    a = _bool_batch(ctx, [True, False, True, None, False])
    assert ctx.boolean_op("and", a, 0, _bool_batch(ctx, [True, True, False, True, None]), 0).download_column(0).to_list() == [True, False, False, None, None]
    assert ctx.boolean_op("or", a, 0, _bool_batch(ctx, [False, True, False, True, None]), 0).download_column(0).to_list() == [True, True, True, None, None]
    assert ctx.boolean_op("not", _bool_batch(ctx, [True, False, None, True]), 0).download_column(0).to_list() == [False, True, None, False]
    c = _bool_batch(ctx, [True, False, True, None, False, True])
    assert c.count_true(0) == 3                                            # nulls count in neither (boolean.rs:669-682)
    assert ctx.boolean_op("not", c, 0).count_true(0) == 2                  # count_false
    with pytest.raises(capi.RivulusError, match="Array lengths must match for logical operations"):
        ctx.boolean_op("and", _bool_batch(ctx, [True, False]), 0, _bool_batch(ctx, [True]), 0)


def test_boolean_ops_random_views_and_compound_predicate(ctx):
    """Strict-null and/or/not over bit-offset views, checked against the definition (value(i) pairs, boolean.rs:127-132), then used
    as a compound predicate: (k > 300) AND NOT (x < 500.0) -> mask column -> RecordBatch::filter."""
    rng = np.random.default_rng(21)
    n = 100_003
    ca, cb = random_col(rng, "bool", n, 0.2, offset=5, tail=3), random_col(rng, "bool", n, 0.1, offset=37)
    ga, gb = upload(ctx, [ca]), upload(ctx, [cb])

    def logical(c):
        vals = c.values[c.offset:c.offset + n].astype(bool)
        valid = np.ones(n, bool) if c.valid is None else c.valid[c.offset:c.offset + n]
        return vals, valid

    (av, ak), (bv, bk) = logical(ca), logical(cb)
    for op, fn in (("and", np.logical_and), ("or", np.logical_or)):
        got = ctx.boolean_op(op, ga, 0, gb, 0).download_column(0)
        valid = ak & bk
        want_vals = fn(av, bv) & valid
        assert np.array_equal(capi.unpack_bits(got.values, n), want_vals), op
        assert got.null_count == int((~valid).sum()) and np.array_equal(capi.unpack_bits(got.validity, n), valid), op
    got = ctx.boolean_op("not", ga, 0).download_column(0)
    assert np.array_equal(capi.unpack_bits(got.values, n), ~av & ak) and np.array_equal(capi.unpack_bits(got.validity, n), ak)
    nonull = ctx.boolean_op("and", upload(ctx, [Col("bool", 70, np.ones(70, bool))]), 0, upload(ctx, [Col("bool", 70, np.zeros(70, bool))]), 0)
    assert nonull.view(0).validity is None and nonull.count_true(0) == 0   # bitmap only when a result is null (boolean.rs:280-286)

    # compound predicate through masks
    cols = [random_col(rng, "i64", n, 0.1, lo=0, hi=1000), random_col(rng, "f64", n, 0.1), random_col(rng, "str", n, 0.1, maxlen=9)]
    batch = upload(ctx, cols)
    m1 = ctx.predicate_mask(batch, capi.predicate(0, ">", 300))
    m2 = ctx.predicate_mask(batch, capi.predicate(1, "<", 500.0))
    mask = ctx.boolean_op("and", m1, 0, ctx.boolean_op("not", m2, 0), 0)
    # expected keep set from the eager truth table: nulls pass `<` (so NOT(x < 500) drops them), nulls fail `>`
    kv = cols[0].values[:n]; kok = cols[0].valid[:n]
    xv = cols[1].values[:n]; xok = cols[1].valid[:n]
    keep = (kok & (kv > 300)) & ~(~xok | (xv < 500.0))
    joined = ctx.upload([c.gpu() for c in cols] + [capi.Column(capi.BOOLEAN, n, 0, capi.pack_bits(keep))])
    want = ctx.filter_project(joined, capi.mask_predicate(3), [0, 1, 2])
    views = [batch.view(i) for i in range(3)] + [mask.view(0)]
    got = ctx.filter_project(ctx.wrap_device(views), capi.mask_predicate(3), [0, 1, 2])
    assert got.num_rows() == want.num_rows() == int(keep.sum())
    for j in range(3):
        assert got.checksum(j) == want.checksum(j)



@pytest.mark.gpu
@pytest.mark.parametrize("keytype", ["i64", "str"])
def test_hash_join_inner_large(keytype):
    """rvl_hash_join_inner through the C ABI at a size the Python model still handles: 300 K build rows x 500 K probe rows, skewed keys
    with nulls on both sides; pairs must come out in probe order, build order within a probe row (plan.rs:198-205)."""
    rng = np.random.default_rng(77)
    nb, npr = 300_000, 500_000
    ctx = capi.Context(0)
    bkeys = rng.integers(0, 200_000, nb).astype(np.int64)
    pkeys = rng.integers(0, 220_000, npr).astype(np.int64)
    bkeys[rng.random(nb) < 0.05] = 7            # a heavy key
    bvalid, pvalid = rng.random(nb) > 0.02, rng.random(npr) > 0.02
    bpay, ppay = np.arange(nb, dtype=np.int64), np.arange(npr, dtype=np.int64) * 3
    if keytype == "i64":
        bcol = capi.Column(capi.INT64, nb, 0, bkeys, capi.pack_bits(bvalid))
        pcol = capi.Column(capi.INT64, npr, 0, pkeys, capi.pack_bits(pvalid))
    else:
        bcol = capi.Column.from_list([("k%d" % k) if v else None for k, v in zip(bkeys, bvalid)], capi.STRING)
        pcol = capi.Column.from_list([("k%d" % k) if v else None for k, v in zip(pkeys, pvalid)], capi.STRING)
    build = ctx.upload([bcol, capi.Column(capi.INT64, nb, 0, bpay)])
    probe = ctx.upload([pcol, capi.Column(capi.INT64, npr, 0, ppay)])
    before = ctx.launch_count()
    out = ctx.hash_join_inner(build, 0, probe, 0, [1], [1])
    assert ctx.launch_count() > before
    got = out.download()
    # model: dict of build rows per key (None = null key), probe order
    table = {}
    for i in range(nb):
        table.setdefault(int(bkeys[i]) if bvalid[i] else None, []).append(i)
    wp, wb = [], []
    for i in range(npr):
        for b in table.get(int(pkeys[i]) if pvalid[i] else None, ()):
            wp.append(i); wb.append(b)
    assert out.num_rows() == len(wp)
    assert np.array_equal(got[0].values, ppay[np.array(wp)]) and np.array_equal(got[1].values, bpay[np.array(wb)])


@pytest.mark.gpu
def test_large_blocks_are_recycled_by_the_context():
    """runtime.cu: blocks >= 16 MiB go back to the context's own free list, not to the driver pool (which was measured re-creating
    multi-GB blocks after a synchronisation: profiles/r02_alloc_outlier.txt).  Steady-state queries must not change what the pool
    holds, also across device-wide synchronisations and small queries in between; rvl_ctx_trim gives the cache back."""
    ctx = capi.Context(0)
    t = ctx.gen_batch([(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_I64, 1, 0), (capi.SYNTH_F64, 2, 0)], 40_000_000, 0)
    small = t.slice(0, 3_000_000)

    def q(b, thr):
        o = ctx.filter_project(b, capi.predicate(0, ">", thr), [1, 2]); n = o.num_rows(); o.release()
        return n
    for thr in (998, 499, 99):
        q(t, thr)
    q(small, 499)
    reserved0, used0 = ctx.pool_stats()
    for rep in range(3):
        for thr in (998, 499, 99):
            q(t, thr)
        q(small, 499)
        ctx.synchronize()
        assert ctx.pool_stats()[0] == reserved0, "a steady-state query made the pool grow"
    assert ctx.pool_stats()[1] >= used0 - (1 << 20)      # the cached blocks still count as used by the pool
    small.release(); t.release()
    ctx.trim()
    assert ctx.pool_stats()[0] < reserved0, "trim did not give the cached blocks back"
