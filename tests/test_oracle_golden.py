"""Pins the CPU oracle against the reference's OWN known-answer tests.

Every test names the reference test it ports as file:line under /root/reference/src.  The
reference cannot be run here (no rustc), so these known-answer pairs are the anchor for parity;
the GPU parity tests then compare the CUDA path with this oracle.
"""
import math

import numpy as np
import pytest

from oracle import oracle as O
from oracle.oracle import (Array, DataFrame, LazyFrame, OracleError, RecordBatch, RecordBatchBuilder, StreamingPhysicalPlan, col, lit, set_extensions,
                           set_csv_reference_validity, calculate_adaptive_batch_size, dtype_is_numeric, dtype_is_comparable_with,
                           EX_BOOLEAN, EX_FLOAT64, EX_INT64, EX_NULL, EX_STRING)


# ---------------------------------------------------------------- fixtures (reference builders)
def df_name_age_score():
    # physical_plan/plan.rs:295-327, logical_plan/builder.rs:128-160
    return DataFrame.new([("name", ["Alice", "Bob", "Charlie"]), ("age", [25, 30, 35]), ("score", [85.5, 92.0, 78.5])])


def df_name_age_active():
    # physical_plan/streaming_planner.rs:177-209
    return DataFrame.new([("name", ["Alice", "Bob", "Charlie"]), ("age", [25, 30, 35]), ("active", [True, False, True])])


def rb_id_name_active():
    # execution/record_batch.rs:585-604 (3 rows, one null string)
    return RecordBatch.try_new(["id", "name", "active"],
                               [Array.from_list([1, 2, 3], EX_INT64), Array.from_list(["Alice", None, "Charlie"], EX_STRING),
                                Array.from_list([True, False, True], EX_BOOLEAN)])


def make_batch(id_start):
    # execution/stream.rs:236-250, physical_plan/streaming.rs:375-389
    return RecordBatch.try_new(["id", "name", "active"],
                               [Array.from_list([id_start, id_start + 1], EX_INT64),
                                Array.from_list([f"name_{id_start}", f"name_{id_start + 1}"], EX_STRING),
                                Array.from_list([True, False], EX_BOOLEAN)])


def pattern(n):
    # execution/array/bitmap.rs:207-210
    return [((i % 3 == 0) != (i % 5 == 0)) for i in range(n)]


# ---------------------------------------------------------------- datatypes/series.rs
def test_anyvalue_partial_ord():  # series.rs:349-366
    assert O.any_partial_cmp(1, 2) == -1
    assert O.any_partial_cmp(1.0, 2.0) == -1
    assert O.any_partial_cmp("a", "b") == -1
    assert O.any_partial_cmp(False, True) == -1
    assert O.any_partial_cmp(None, 0) == -1          # Null < everything
    assert O.any_partial_cmp(None, False) == -1
    assert O.any_partial_cmp(1, "1") is None         # cross-type => None


def test_anyvalue_eq_semantics():  # series.rs:87-98 (+ tests :310-335)
    assert O.any_eq(None, None)
    assert O.any_eq(42, 42) and not O.any_eq(42, 43)
    assert not O.any_eq(1, 1.0)                      # no Int<->Float coercion
    assert O.any_eq(0.0, -0.0)
    assert not O.any_eq(float("nan"), float("nan"))


def test_series_dtype_inference():  # series.rs:400-469, 521-534
    assert DataFrame.new([("numbers", [1, 2, 3])]).dtypes() == ["Int64"]
    assert DataFrame.new([("letters", ["a", "b", "c"])]).dtypes() == ["String"]
    assert DataFrame.new([("with_nulls", [1, None, 3])]).dtypes() == ["Int64"]
    assert DataFrame.new([("nulls", [None, None])]).dtypes() == ["Null"]
    assert DataFrame.new([("mixed_numeric", [1, 2.5, 3])]).dtypes() == ["Float64"]
    with pytest.raises(OracleError, match="Mixed types in series: expected Int64, found String"):
        DataFrame.new([("mixed", [1, "hello"])])
    with pytest.raises(OracleError, match="Empty series not allowed"):
        DataFrame.new([("empty", [])])


# ---------------------------------------------------------------- physical_plan/plan.rs (eager)
def test_dataframe_construction_rules():  # datatypes/dataframe.rs tests: creation :41-63, length mismatch :84-118, duplicate :120-150, edge cases :end
    df = df_name_age_score()
    assert (df.height(), df.width(), df.column_names()) == (3, 3, ["name", "age", "score"])
    e = DataFrame.new([])                                                     # test_dataframe_creation_empty_columns_list
    assert (e.height(), e.width()) == (0, 0)
    one = DataFrame.new([("name", ["Alice", "Bob"])])                         # test_dataframe_creation_single_column
    assert (one.height(), one.width()) == (2, 1)
    with pytest.raises(OracleError, match="Column lengths mismatch: expected 2, found 3 for column 'age'"):
        DataFrame.new([("name", ["Alice", "Bob"]), ("age", [25, 30, 35])])
    with pytest.raises(OracleError, match="Duplicate column name: 'name'"):
        DataFrame.new([("name", ["Alice", "Bob"]), ("name", ["Charlie", "David"])])
    nulls = DataFrame.new([("nulls", [None, None, None]), ("numbers", [1, 2, 3])])   # test_dataframe_with_all_null_column
    assert (nulls.height(), nulls.width(), nulls.dtypes()) == (3, 2, ["Null", "Int64"])
    mixed = DataFrame.new([("mixed", [1, 2.5, 3])])                           # test_dataframe_with_mixed_numeric_column
    assert mixed.dtypes() == ["Float64"]


def test_execute_filter_gt():  # plan.rs:505-525
    r = LazyFrame.from_dataframe(df_name_age_score()).filter(col("age").gt(lit(25))).collect()
    assert (r.height(), r.width()) == (2, 3)
    assert r.column("age") == [30, 35]
    assert r.column("name") == ["Bob", "Charlie"]


def test_execute_filter_eq():  # plan.rs:528-547
    r = LazyFrame.from_dataframe(df_name_age_score()).filter(col("name").eq(lit("Bob"))).collect()
    assert r.height() == 1 and r.column("name") == ["Bob"]


def test_execute_filter_lt():  # plan.rs:550-569
    r = LazyFrame.from_dataframe(df_name_age_score()).filter(col("score").lt(lit(90.0))).collect()
    assert r.height() == 2 and r.column("name") == ["Alice", "Charlie"]


def test_execute_filter_no_matches():  # plan.rs:572-589
    r = LazyFrame.from_dataframe(df_name_age_score()).filter(col("age").gt(lit(100))).collect()
    assert (r.height(), r.width()) == (0, 3)
    assert r.column_names() == ["name", "age", "score"]
    assert r.dtypes() == ["String", "Int64", "Float64"]     # Series::empty keeps the dtype (plan.rs:140-141)


def test_execute_limit():  # plan.rs:615-667
    lf = LazyFrame.from_dataframe(df_name_age_score())
    r = lf.limit(2).collect()
    assert (r.height(), r.width()) == (2, 3) and r.column("name") == ["Alice", "Bob"]
    r = lf.limit(10).collect()
    assert r.height() == 3
    r = lf.limit(0).collect()
    assert (r.height(), r.width()) == (0, 3)


def test_execute_chained_operations():  # plan.rs:672-704
    r = (LazyFrame.from_dataframe(df_name_age_score()).select([col("name"), col("age"), col("score")])
         .filter(col("age").gt(lit(25))).limit(1).collect())
    assert (r.height(), r.width()) == (1, 3) and r.column("name") == ["Bob"]


def test_execute_filter_then_select():  # plan.rs:707-735
    r = (LazyFrame.from_dataframe(df_name_age_score()).filter(col("age").gte(lit(30)))
         .select([col("name"), col("score")]).collect())
    assert (r.height(), r.width()) == (2, 2)
    assert r.column_names() == ["name", "score"] and r.column("name") == ["Bob", "Charlie"]
    assert r.column("score") == [92.0, 78.5]


def test_execute_select_variants():  # plan.rs:424-502
    lf = LazyFrame.from_dataframe(df_name_age_score())
    assert lf.select([col("name")]).collect().column_names() == ["name"]
    assert lf.select([col("score"), col("name")]).collect().column_names() == ["score", "name"]
    with pytest.raises(OracleError, match="Column not found: 'nonexistent'"):
        lf.select([col("nonexistent")]).collect()


# ---------------------------------------------------------------- logical_plan/builder.rs (end to end)
def test_collect_simple_select_filter_limit():  # builder.rs:436-473
    df = df_name_age_score()
    r = LazyFrame.from_dataframe(df).select([col("name"), col("age")]).collect()
    assert (r.width(), r.height()) == (2, 3) and r.column_names() == ["name", "age"]
    r = LazyFrame.from_dataframe(df).filter(col("age").gt(lit(25))).collect()
    assert (r.height(), r.width()) == (2, 3)
    r = LazyFrame.from_dataframe(df).limit(2).collect()
    assert (r.height(), r.width()) == (2, 3)


def test_collect_invalid_columns():  # builder.rs:497-532
    with pytest.raises(OracleError, match=r"Logical plan error: Column not found: 'nonexistent'"):
        LazyFrame.from_dataframe(df_name_age_score()).select([col("nonexistent")]).collect()
    with pytest.raises(OracleError, match=r"Logical plan error: Column not found: 'nonexistent'"):
        LazyFrame.from_dataframe(df_name_age_score()).filter(col("nonexistent").gt(lit(0))).collect()


def test_collect_streaming():  # builder.rs:570-614
    df = df_name_age_score()
    r = LazyFrame.from_dataframe(df).select([col("name"), col("age")]).collect_streaming()
    assert (r.num_columns(), r.num_rows()) == (2, 3) and r.column_names() == ["name", "age"]
    with pytest.raises(OracleError):          # no 'active' column in this frame (builder.rs:584-594)
        LazyFrame.from_dataframe(df).filter(col("active")).collect_streaming()
    old = LazyFrame.from_dataframe(df).select([col("name")]).collect()
    new = LazyFrame.from_dataframe(df).select([col("name")]).collect_streaming()
    assert (old.width(), old.height()) == (new.num_columns(), new.num_rows())


def test_readme_shape_fails_validation():  # README.md:59-63 vs logical_plan/plan.rs:139-146
    lf = LazyFrame.from_dataframe(df_name_age_score()).select([col("name")]).filter(col("age").gt(lit(25)))
    assert lf.plan_shape() == "Filter(Select(Source))"
    with pytest.raises(OracleError, match="Logical plan error: Column not found: 'age'"):
        lf.collect()


def test_optimizer_rewrite():  # optimizer.rs:15-64
    df = df_name_age_score()
    # predicate column among the selected ones => Select(Filter) becomes Filter(Select)
    lf = LazyFrame.from_dataframe(df).filter(col("age").gt(lit(25))).select([col("name"), col("age")])
    assert lf.plan_shape() == "Filter(Select(Source))"
    assert lf.collect().to_dict() == {"name": ["Bob", "Charlie"], "age": [30, 35]}
    # predicate column not selected => unchanged
    lf = LazyFrame.from_dataframe(df).filter(col("age").gt(lit(25))).select([col("name")])
    assert lf.plan_shape() == "Select(Filter(Source))"
    assert lf.collect().to_dict() == {"name": ["Bob", "Charlie"]}
    # Limit blocks the recursion (optimizer.rs:62)
    lf = LazyFrame.from_dataframe(df).filter(col("age").lt(lit(40))).select([col("age")]).limit(2)
    assert lf.plan_shape() == "Limit(Select(Filter(Source)))"


def test_planner_rejections():  # planner.rs:134-189 (tests :191-697 duplicate plan.rs's)
    lf = LazyFrame.from_dataframe(df_name_age_score())
    with pytest.raises(OracleError, match="Unsupported filter: only simple column comparisons supported"):
        lf.filter(col("age").gt(lit(1)).and_(col("age").lt(lit(9)))).collect()
    with pytest.raises(OracleError, match="Unsupported binary operator in filter: Plus"):
        lf.filter(col("age").add(lit(1))).collect()
    with pytest.raises(OracleError, match="Filter must be a binary comparison, found: Column"):
        lf.filter(col("age")).collect()
    with pytest.raises(OracleError, match="Filter right side must be a literal value"):
        lf.filter(col("age").gt(col("score"))).collect()
    with pytest.raises(OracleError, match="Unsupported expression"):
        lf.select([col("age").add(lit(1))]).collect()


# ---------------------------------------------------------------- main.rs demo queries
def df_main():
    return DataFrame.new([("name", ["Alice", "Bob", "Charlie", "Diana", "Eve"]), ("age", [25, 30, 35, 28, 42]),
                          ("score", [85.5, 92.0, 78.5, 94.5, 88.0])])


def test_main_demo_queries():  # main.rs:47-94
    df = df_main()
    q1 = LazyFrame.from_dataframe(df).select([col("name"), col("age")]).filter(col("age").gt(lit(30))).collect()
    assert q1.to_dict() == {"name": ["Charlie", "Eve"], "age": [35, 42]}
    q2 = (LazyFrame.from_dataframe(df).filter(col("score").gte(lit(90.0)))
          .select([col("name"), col("age").alias("user_age")]).collect())
    assert q2.to_dict() == {"name": ["Bob", "Diana"], "user_age": [30, 28]}
    q3 = LazyFrame.from_dataframe(df).filter(col("age").lt(lit(40))).limit(2).collect()
    assert q3.to_dict() == {"name": ["Alice", "Bob"], "age": [25, 30], "score": [85.5, 92.0]}
    q4 = LazyFrame.from_dataframe(df).filter(col("age").gt(lit(100))).collect()
    assert (q4.height(), q4.column_names(), q4.dtypes()) == (0, ["name", "age", "score"], ["String", "Int64", "Float64"])
    q5 = LazyFrame.from_dataframe(df).select([col("name")]).limit(0).collect()
    assert (q5.height(), q5.column_names(), q5.dtypes()) == (0, ["name"], ["String"])


# ---------------------------------------------------------------- execution/record_batch.rs
def test_rb_slice():  # record_batch.rs:700-749
    b = rb_id_name_active()
    s = b.slice(1, 2)
    assert (s.num_rows(), s.num_columns()) == (2, 3)
    assert s.column(0).to_list() == [2, 3]
    assert b.slice(1, 0).num_rows() == 0 and b.slice(0, 3).num_rows() == 3 and b.slice(2, 1).num_rows() == 1
    with pytest.raises(OracleError) as ei:
        b.slice(2, 5)
    assert ei.value.panic and "Slice out of bounds" in str(ei.value)


def test_rb_take():  # record_batch.rs:752-791
    b = rb_id_name_active()
    t = b.take([2, 0, 1])
    assert t.column(0).to_list() == [3, 1, 2]
    assert t.column(1).to_list() == ["Charlie", "Alice", None]
    assert b.take([]).num_rows() == 0
    with pytest.raises(OracleError, match="Index 5 out of bounds for 3 rows"):
        b.take([0, 5, 1])


def test_rb_mixed_types_null_column_and_chained_ops():  # record_batch.rs:1049-1073, 1076-1103, 1106-1130, 1133-1151
    # test_mixed_column_types: one column of every array type, a null string among them
    m = RecordBatch.try_new(["int_col", "float_col", "string_col", "bool_col", "null_col"],
                            [Array.from_list([1, 2, 3], EX_INT64), Array.from_list([1.1, 2.2, 3.3], EX_FLOAT64),
                             Array.from_list(["A", None, "C"], EX_STRING), Array.from_list([True, False, True], EX_BOOLEAN),
                             Array.from_list([None, None, None], EX_NULL)])
    assert (m.num_rows(), m.num_columns()) == (3, 5)
    # test_null_column_handling: the NullArray column reports dtype Null and null_count == len
    nb = RecordBatch.try_new(["id", "null_col"], [Array.from_list([1, 2, 3], EX_INT64), Array.from_list([None, None, None], EX_NULL)])
    assert (nb.num_rows(), nb.num_columns()) == (3, 2)
    assert nb.column(1).dtype == EX_NULL and nb.column(1).null_count == 3
    # test_large_batch_performance (the assertions, not the timings): 10 000 rows, slice(1000, 5000)
    size = 10_000
    big = RecordBatch.try_new(["id", "value"], [Array.from_list(list(range(size)), EX_INT64), Array.from_list([i * 1.5 for i in range(size)], EX_FLOAT64)])
    assert big.num_rows() == size
    sl = big.slice(1000, 5000)
    assert sl.num_rows() == 5000 and sl.column(0).to_list()[:2] == [1000, 1001] and sl.column(1).to_list()[-1] == 5999 * 1.5
    # test_slice_preserves_schema
    b = rb_id_name_active()
    assert b.slice(1, 1).column_names() == ["id", "name", "active"]
    # test_chained_operations: slice -> select -> filter
    f = b.slice(0, 3).select_columns([0, 2]).filter(Array.from_list([True, False, True], EX_BOOLEAN))
    assert (f.num_rows(), f.num_columns()) == (2, 2) and f.column(0).to_list() == [1, 3] and f.column(1).to_list() == [True, True]


def test_rb_select_columns():  # record_batch.rs:794-819
    b = rb_id_name_active()
    assert b.select_columns([0, 2]).column_names() == ["id", "active"]
    assert b.select_columns_by_name(["name", "id"]).column_names() == ["name", "id"]


def test_rb_filter():  # record_batch.rs:822-879
    b = rb_id_name_active()
    f = b.filter(Array.from_list([True, False, True], EX_BOOLEAN))
    assert (f.num_rows(), f.num_columns()) == (2, 3) and f.column(0).to_list() == [1, 3]
    assert b.filter(Array.from_list([True, True, True], EX_BOOLEAN)).num_rows() == 3
    assert b.filter(Array.from_list([False, False, False], EX_BOOLEAN)).num_rows() == 0
    assert b.filter(Array.from_list([True, None, False], EX_BOOLEAN)).num_rows() == 1   # null mask entry drops the row
    with pytest.raises(OracleError, match="Predicate must be a BooleanArray"):
        b.filter(Array.from_list([1, 2, 3], EX_INT64))
    with pytest.raises(OracleError, match="Predicate length 2 doesn't match batch length 3"):
        b.filter(Array.from_list([True, False], EX_BOOLEAN))


def test_rb_concat():  # record_batch.rs:882-949
    b1 = RecordBatch.try_new(["id", "name", "active"], [Array.from_list([1, 2], EX_INT64), Array.from_list(["A", None], EX_STRING),
                                                        Array.from_list([True, False], EX_BOOLEAN)])
    b2 = RecordBatch.try_new(["id", "name", "active"], [Array.from_list([3, 4], EX_INT64), Array.from_list(["B", "C"], EX_STRING),
                                                        Array.from_list([True, True], EX_BOOLEAN)])
    c = RecordBatch.concat([b1, b2])
    assert (c.num_rows(), c.num_columns()) == (4, 3)
    assert c.column(0).to_list() == [1, 2, 3, 4]
    assert c.column(1).to_list() == ["A", None, "B", "C"]
    e = b1.empty_like()
    assert RecordBatch.concat([e, e]).num_rows() == 0
    other = RecordBatch.try_new(["name"], [Array.from_list([], EX_STRING)])
    idonly = RecordBatch.try_new(["id"], [Array.from_list([], EX_INT64)])
    with pytest.raises(OracleError, match="All batches must have the same schema"):
        RecordBatch.concat([idonly, other])
    with pytest.raises(OracleError, match="Cannot concatenate empty batch list"):
        RecordBatch.concat([])


def test_rb_try_new_errors():  # record_batch.rs:16-58 (tests :620-697)
    with pytest.raises(OracleError, match="Schema has 2 fields but 1 columns provided"):
        RecordBatch.try_new(["a"], [Array.from_list([1], EX_INT64)], schema_dtypes=[EX_INT64, EX_INT64], schema_names=["a", "b"])
    with pytest.raises(OracleError, match="Column 1 has length 2 but expected 3"):
        RecordBatch.try_new(["a", "b"], [Array.from_list([1, 2, 3], EX_INT64), Array.from_list([1, 2], EX_INT64)])
    with pytest.raises(OracleError, match="Column 0 has type Int64 but schema expects String"):
        RecordBatch.try_new(["a"], [Array.from_list([1], EX_INT64)], schema_dtypes=[EX_STRING])


def test_output_layout_rules():
    # primitive.rs:175-185 / record_batch.rs:140-147: placeholder 0 under nulls, bitmap present iff a survivor is null
    b = RecordBatch.try_new(["v", "w"], [Array.from_list([10, None, 30, None], EX_INT64), Array.from_list([1.5, 2.5, None, 4.5], EX_FLOAT64)])
    t = b.take([0, 1, 2])
    v, w = t.column(0), t.column(1)
    assert v.values.tolist() == [10, 0, 30] and v.validity is not None and v.validity.tolist() == [0b101] and v.offset == 0
    assert w.values.tolist() == [1.5, 2.5, 0.0] and w.validity.tolist() == [0b011]
    t2 = b.take([0, 2])
    assert t2.column(0).validity is None and t2.column(0).null_count == 0        # no nulls => bitmap dropped
    assert t2.column(1).validity is not None


# ---------------------------------------------------------------- arrays / bitmaps
def test_bitmap_roundtrip_and_slices():  # bitmap.rs:230-241, 255-309, 336-361
    vals = pattern(37)
    bits = O.pack_bits(vals)
    a = Array.boolean(bits, 37)
    assert a.to_list() == vals
    assert bits[0] == sum((1 << i) for i in range(8) if vals[i])            # LSB-first
    vals = pattern(64)
    a = Array.boolean(O.pack_bits(vals), 64)
    s = a.slice(7, 25)
    assert s.export().offset == 7 and s.to_list() == vals[7:32]
    vals = pattern(91)
    a = Array.boolean(O.pack_bits(vals), 91)
    s1, s2 = a.slice(10, 50).slice(7, 20), a.slice(17, 20)
    assert s1.export().offset == s2.export().offset == 17 and s1.to_list() == s2.to_list() == vals[17:37]
    vals = [i in (7, 8, 16) for i in range(17)]
    a = Array.boolean(O.pack_bits(vals), 17)
    assert a.to_list() == vals and a.slice(6, 6).to_list() == vals[6:12]
    with pytest.raises(OracleError):
        Array.boolean(O.pack_bits([False] * 16), 16).slice(9, 8)


def test_primitive_slice_with_nulls():  # primitive.rs:283-306
    a = Array.from_list([1, None, 3, None, 5, None], EX_INT64)
    s = a.slice(2, 3)
    e = s.export()
    assert e.length == 3 and e.offset == 2
    assert s.to_list() == [3, None, 5] and s.null_count() == 1


def test_primitive_builder():  # primitive.rs:546-604
    b = RecordBatch.try_new(["x"], [Array.from_list([10, None, 20, 30, None], EX_INT64)]).take([0, 1, 2, 3, 4]).column(0)
    assert b.length == 5 and b.null_count == 2 and b.to_list() == [10, None, 20, 30, None]
    assert b.values.tolist() == [10, 0, 20, 30, 0]
    c = RecordBatch.try_new(["x"], [Array.from_list([10, 20, 30, 40, 50], EX_INT64)]).take([0, 1, 2, 3, 4]).column(0)
    assert c.null_count == 0 and c.validity is None


def test_array_slices_and_null_counts():  # boolean.rs / string.rs test_slice_operation, primitive.rs test_multiple_slices_consistency,
    # test_new_with_some_nulls / test_all_nulls / test_empty_array of the three array types
    b = Array.from_list([True, None, False, None, True, False], EX_BOOLEAN).slice(2, 3)
    assert (b.len(), b.export().offset, b.to_list(), b.null_count()) == (3, 2, [False, None, True], 1)
    st = Array.from_list(["a", None, "b", None, "c", "d"], EX_STRING).slice(2, 3)
    assert (st.len(), st.export().offset, st.to_list(), st.null_count()) == (3, 2, ["b", None, "c"], 1)
    vals = list(range(20))
    arr = Array.from_list([v if v % 3 != 0 else None for v in vals], EX_INT64)     # every third is null
    s1, s2, s3 = arr.slice(0, 10), arr.slice(5, 10), arr.slice(10, 10)
    assert s1.null_count() + s3.null_count() == arr.null_count() == 7
    assert s1.to_list()[5:10] == s2.to_list()[0:5]
    for dtype, some, alln in ((EX_INT64, [1, None, 3], [None, None]), (EX_FLOAT64, [1.5, None, 2.5], [None, None]),
                              (EX_BOOLEAN, [True, None, False], [None, None]), (EX_STRING, ["x", None, ""], [None, None])):
        a = Array.from_list(some, dtype)
        assert (a.len(), a.null_count(), a.to_list()) == (3, 1, some)
        n = Array.from_list(alln, dtype)
        assert (n.len(), n.null_count(), n.to_list()) == (2, 2, alln)
        e = Array.from_list([], dtype)
        assert (e.len(), e.null_count(), e.to_list()) == (0, 0, [])


def test_string_layout():  # string.rs:362-381, 574-633, 683-694
    a = Array.from_list(["hello", "", "world"], EX_STRING).export()
    assert a.offsets.tolist() == [0, 5, 5, 10] and a.validity is None
    a = Array.from_list(["hello", None, "world!!!", ""], EX_STRING)
    e = a.export()
    assert e.offsets.tolist() == [0, 5, 5, 13, 13] and a.to_list() == ["hello", None, "world!!!", ""]   # nulls zero-length
    e = Array.from_list(["ascii", "café", "🦀", "🦀🔥", "", None], EX_STRING).export()
    assert np.diff(e.offsets).tolist() == [5, 5, 4, 8, 0, 0]
    e = Array.from_list(["hello", "world", None, "test"], EX_STRING).export()
    assert len(e.data) == 14                                                                           # total_bytes
    assert Array.from_list([], EX_STRING).export().validity is None


# ---------------------------------------------------------------- streams (stream.rs, streaming.rs)
def test_streaming_plan_memory_source_ops():  # streaming.rs:392-462
    b1, b2 = make_batch(1), make_batch(3)
    r = StreamingPhysicalPlan.memory_source([b1, b2]).collect()
    assert (r.num_rows(), r.num_columns()) == (4, 3)
    r = StreamingPhysicalPlan.memory_source([b1, b2]).filter("active").collect()
    assert (r.num_rows(), r.num_columns()) == (2, 3) and r.column(0).to_list() == [1, 3]
    r = StreamingPhysicalPlan.memory_source([b1, b2]).select(["id", "name"]).collect()
    assert (r.num_rows(), r.num_columns()) == (4, 2) and r.column_names() == ["id", "name"]
    r = StreamingPhysicalPlan.memory_source([b1, b2]).limit(3).collect()
    assert (r.num_rows(), r.num_columns()) == (3, 3)
    r = StreamingPhysicalPlan.memory_source([b1, b2]).filter("active").select(["name"]).limit(1).collect()
    assert (r.num_rows(), r.num_columns()) == (1, 1) and r.column_names() == ["name"] and r.column(0).to_list() == ["name_1"]


def test_limit_stream_batches():  # streaming.rs:465-498
    b1, b2 = make_batch(1), make_batch(3)
    got = StreamingPhysicalPlan.memory_source([b1, b2]).limit(2).collect_batches()
    assert [b.num_rows() for b in got] == [2]                     # exact batch size: second pull returns None
    got = StreamingPhysicalPlan.memory_source([b1, b2]).limit(3).collect_batches()
    assert [b.num_rows() for b in got] == [2, 1]                  # partial batch, then None


def test_collect_vs_collect_batches():  # streaming.rs:501-516, stream.rs:535-550
    b1, b2 = make_batch(1), make_batch(3)
    assert StreamingPhysicalPlan.memory_source([b1, b2]).collect().num_rows() == 4
    assert [b.num_rows() for b in StreamingPhysicalPlan.memory_source([b1, b2]).collect_batches()] == [2, 2]


def test_filter_select_stream_operators():  # stream.rs:366-432, 437-532
    b = make_batch(1)
    got = StreamingPhysicalPlan.memory_source([b]).filter("active").collect_batches()
    assert [x.num_rows() for x in got] == [1]
    none = RecordBatch.try_new(["id", "name", "active"], [Array.from_list([1, 2], EX_INT64), Array.from_list(["a", "b"], EX_STRING),
                                                         Array.from_list([False, False], EX_BOOLEAN)])
    got = StreamingPhysicalPlan.memory_source([none]).filter("active").collect_batches()
    assert [x.num_rows() for x in got] == [0]
    with pytest.raises(OracleError, match="Column 'nonexistent' not found in schema"):
        StreamingPhysicalPlan.memory_source([b]).filter("nonexistent").collect()
    with pytest.raises(OracleError, match="Predicate column 'id' is not of boolean type"):
        StreamingPhysicalPlan.memory_source([b]).filter("id").collect()
    with pytest.raises(OracleError, match="Column 'nonexistent' not found in schema"):
        StreamingPhysicalPlan.memory_source([b]).select(["nonexistent"]).collect()
    r = StreamingPhysicalPlan.memory_source([b]).filter("active").select(["name"]).collect_batches()[0]
    assert (r.num_columns(), r.num_rows(), r.column_names()) == (1, 1, ["name"])
    with pytest.raises(OracleError, match="Cannot create stream from empty batch list"):
        StreamingPhysicalPlan.memory_source([]).collect()


# ---------------------------------------------------------------- streaming_planner.rs (from DataFrame)
def test_streaming_planner_conversions():  # streaming_planner.rs:212-329
    df = df_name_age_active()
    r = LazyFrame.from_dataframe(df).collect_streaming()
    assert (r.num_rows(), r.num_columns()) == (3, 3)
    r = LazyFrame.from_dataframe(df).select([col("name"), col("age")]).collect_streaming()
    assert r.column_names() == ["name", "age"]
    r = LazyFrame.from_dataframe(df).filter(col("active")).collect_streaming()
    assert r.num_rows() == 2 and r.column(0).to_list() == ["Alice", "Charlie"]
    r = LazyFrame.from_dataframe(df).limit(2).collect_streaming()
    assert (r.num_rows(), r.num_columns()) == (2, 3)
    r = LazyFrame.from_dataframe(df).filter(col("active")).select([col("name")]).limit(1).collect_streaming()
    assert (r.num_rows(), r.num_columns(), r.column_names()) == (1, 1, ["name"])


def test_streaming_planner_rejections():  # streaming_planner.rs:332-381
    df = df_name_age_active()
    with pytest.raises(OracleError, match="Expression conversion error"):
        LazyFrame.from_dataframe(df).select([col("age").add(lit(10))]).collect_streaming()
    with pytest.raises(OracleError, match="Binary expressions not yet supported"):
        LazyFrame.from_dataframe(df).filter(col("age").gt(lit(30))).collect_streaming()


def test_streaming_alias_dropped_and_null_flattening():  # streaming_planner.rs:110-113; streaming.rs:177,188,212
    df = DataFrame.new([("name", ["a", None, "c"]), ("age", [1, None, 3]), ("x", [1.5, None, 2.5]), ("ok", [True, None, False])])
    r = LazyFrame.from_dataframe(df).select([col("name"), col("age").alias("years"), col("x"), col("ok")]).collect_streaming()
    assert r.column_names() == ["name", "age", "x", "ok"]
    assert r.to_dict() == {"name": ["a", None, "c"], "age": [1, 0, 3], "x": [1.5, 0.0, 2.5], "ok": [True, False, False]}
    assert r.column(1).validity is None and r.column(0).validity is not None


def test_streaming_batches_of_1024():  # streaming_planner.rs:32 + streaming.rs:144-167
    n = 2500
    df = DataFrame.new([("i", list(range(n))), ("flag", [(i % 3 == 0) for i in range(n)])])
    r = LazyFrame.from_dataframe(df).filter(col("flag")).limit(700).collect_streaming()
    assert r.num_rows() == 700 and r.column(0).to_list() == [i for i in range(n) if i % 3 == 0][:700]


# ---------------------------------------------------------------- not pinned by a reference test ("code reading")
S1 = {  # (row, literal) -> {op: expected}
    "null_vs_value": ((None, 30), {"==": False, "!=": True, "<": True, "<=": True, ">": False, ">=": False}),
    "value_vs_null": ((30, None), {"==": False, "!=": True, "<": False, "<=": False, ">": True, ">=": True}),
    "null_vs_null": ((None, None), {"==": True, "!=": False, "<": False, "<=": True, ">": False, ">=": True}),
    "nan_row": ((float("nan"), 1.0), {"==": False, "!=": True, "<": False, "<=": False, ">": False, ">=": False}),
    "nan_lit": ((1.0, float("nan")), {"==": False, "!=": True, "<": False, "<=": False, ">": False, ">=": False}),
    "cross_type": ((30, 30.0), {"==": False, "!=": True, "<": False, "<=": False, ">": False, ">=": False}),
    "neg_zero": ((-0.0, 0.0), {"==": True, "!=": False, "<": False, "<=": True, ">": False, ">=": True}),
    "bool": ((False, True), {"==": False, "!=": True, "<": True, "<=": True, ">": False, ">=": False}),
    "str_prefix": (("ab", "abc"), {"==": False, "!=": True, "<": True, "<=": True, ">": False, ">=": False}),
}


@pytest.mark.parametrize("case", sorted(S1))
def test_truth_table_unpinned(case):  # plan.rs:114-120 over series.rs:87-117
    (row, literal), exp = S1[case]
    for op, want in exp.items():
        assert O.eval_cmp(row, op, literal) is want, (case, op)


def test_eager_nulls_dtype_collapse_and_empty_errors():  # SURVEY S2/S3 (plan.rs:140-143, 89-91, 167-169)
    df = DataFrame.new([("age", [10, None, 30, None]), ("v", [None, None, 7, None])])
    r = LazyFrame.from_dataframe(df).filter(col("age").lt(lit(20))).collect()     # null ages pass `<`
    assert r.to_dict() == {"age": [10, None, None], "v": [None, None, None]}
    assert r.dtypes() == ["Int64", "Null"]                                        # all-null survivors collapse to Null
    with pytest.raises(OracleError, match="Execution error: Series error: Empty series not allowed"):
        LazyFrame.from_dataframe(df).filter(col("age").gt(lit(100))).select([col("v")]).collect()
    # ...but when the optimizer swaps Select below Filter (predicate column selected) the Select sees rows: no error
    assert LazyFrame.from_dataframe(df).filter(col("age").gt(lit(100))).select([col("age")]).collect().height() == 0
    with pytest.raises(OracleError, match="Execution error: Series error: Empty series not allowed"):
        LazyFrame.from_dataframe(df).filter(col("age").gt(lit(100))).limit(5).collect()
    assert LazyFrame.from_dataframe(df).filter(col("age").gt(lit(100))).limit(0).collect().height() == 0


def test_fused_oracle_matches_eager_on_columnar_input():
    # the fused-operator oracle (used to check the GPU kernel) agrees with the eager engine on the same data
    rng = np.random.default_rng(7)
    n = 300
    k = [None if rng.random() < 0.2 else int(rng.integers(0, 50)) for _ in range(n)]
    x = [None if rng.random() < 0.2 else float(rng.random()) for _ in range(n)]
    s = [None if rng.random() < 0.2 else "s%d" % rng.integers(0, 99) for _ in range(n)]
    df = DataFrame.new([("k", k), ("x", x), ("s", s)])
    rb = RecordBatch.try_new(["k", "x", "s"], [Array.from_list(k, EX_INT64), Array.from_list(x, EX_FLOAT64), Array.from_list(s, EX_STRING)])
    for op in ["==", "!=", "<", ">", "<=", ">="]:
        for literal in (25, None, 25.0):
            eager = LazyFrame.from_dataframe(df).filter(getattr(col("k"), {"==": "eq", "!=": "neq", "<": "lt", ">": "gt", "<=": "lte", ">=": "gte"}[op])(lit(literal))).collect()
            fused = rb.filter_project_cmp(0, op, literal, [0, 1, 2])
            assert fused.num_rows() == eager.height()
            if eager.height():
                assert fused.to_dict() == eager.to_dict(), (op, literal)


# ---------------------------------------------------------------- opt-in extension: no reference behaviour exists
def test_extension_compound_and_streaming_comparison_predicates():
    """And / Or over `column <op> literal` leaves in collect() / collect_streaming(), comparisons in collect_streaming().
    OFF by default: the reference's rejections (planner.rs:146-150, streaming_planner.rs:137-168) are what a user sees.
    ON: every leaf is the eager truth table (plan.rs:114-120; a Null row is Less than any literal, series.rs:105-107), And / Or
    combine the leaves' true / false.  The expected rows below are worked out by hand from that definition."""
    df = DataFrame.new([("name", ["Alice", None, "Charlie", "Diana", "Eve", "Frank"]), ("age", [25, None, 35, 28, 42, 19]),
                        ("score", [85.5, 92.0, None, 94.5, 88.0, 70.0])])
    both = col("age").gt(lit(26)).and_(col("score").lt(lit(93.0)))
    with pytest.raises(OracleError, match="Unsupported filter: only simple column comparisons supported"):
        LazyFrame.from_dataframe(df).filter(both).collect()
    with pytest.raises(OracleError, match="Binary expressions not yet supported in streaming mode"):
        LazyFrame.from_dataframe(df).filter(col("age").gt(lit(26))).collect_streaming()
    set_extensions(True)
    try:
        # eager.  age > 26: F F(null) T T T F;  score < 93: T T T(null is Less) F T T  => rows 2, 4
        out = LazyFrame.from_dataframe(df).filter(both).collect()
        assert out.to_dict() == {"name": ["Charlie", "Eve"], "age": [35, 42], "score": [None, 88.0]}
        # age < 20: F T(null) F F F T;  score >= 94: F F F(null) T F F  => rows 1, 3, 5; select + limit on top
        either = col("age").lt(lit(20)).or_(col("score").gte(lit(94.0)))
        out = LazyFrame.from_dataframe(df).filter(either).select([col("name"), col("score").alias("s")]).collect()
        assert out.to_dict() == {"name": [None, "Diana", "Frank"], "s": [92.0, 94.5, 70.0]}
        assert LazyFrame.from_dataframe(df).filter(either).limit(2).collect().to_dict() == {"name": [None, "Diana"], "age": [None, 28], "score": [92.0, 94.5]}
        # nested: (age > 26 AND score < 93) OR name == "Frank"; cross-type leaf (age == "x") is never true, != always
        nested = both.or_(col("name").eq(lit("Frank")))
        assert LazyFrame.from_dataframe(df).filter(nested).select([col("name")]).collect().to_dict() == {"name": ["Charlie", "Eve", "Frank"]}
        assert LazyFrame.from_dataframe(df).filter(col("age").neq(lit("x")).and_(col("age").gte(lit(35)))).select([col("age")]).collect().to_dict() == {"age": [35, 42]}
        # eager vs streaming on nulls: the eager engine sees Null (< -1 holds), the streaming engine sees the flattened 0 (streaming.rs:177)
        q = col("age").lt(lit(-1)).or_(col("name").eq(lit("Eve")))
        assert LazyFrame.from_dataframe(df).filter(q).select([col("name")]).collect().to_dict() == {"name": [None, "Eve"]}
        assert LazyFrame.from_dataframe(df).filter(q).select([col("name")]).collect_streaming().to_dict() == {"name": ["Eve"]}
        # streaming: a plain comparison (fused device pipeline), a compound one, strings keep their nulls (a null name is Less than "B")
        s1 = LazyFrame.from_dataframe(df).filter(col("age").gt(lit(26))).select([col("name"), col("age")]).collect_streaming()
        assert s1.to_dict() == {"name": ["Charlie", "Diana", "Eve"], "age": [35, 28, 42]}
        s2 = LazyFrame.from_dataframe(df).filter(both).collect_streaming()
        assert s2.to_dict() == {"name": ["Charlie", "Eve"], "age": [35, 42], "score": [0.0, 88.0]}
        s3 = LazyFrame.from_dataframe(df).filter(col("name").lt(lit("B")).and_(col("score").gt(lit(80.0)))).select([col("score")]).limit(1).collect_streaming()
        assert s3.to_dict() == {"score": [85.5]}
        s4 = LazyFrame.from_dataframe(df).filter(col("name").lt(lit("B"))).select([col("age")]).collect_streaming()
        assert s4.to_dict() == {"age": [25, 0]}
        # malformed leaves raise the planner's own messages; unknown columns fail validation as ever
        with pytest.raises(OracleError, match="Filter right side must be a literal value"):
            LazyFrame.from_dataframe(df).filter(both.and_(col("age").gt(col("score")))).collect()
        with pytest.raises(OracleError, match="Streaming planner error: Expression conversion error: Filter right side must be a literal value"):
            LazyFrame.from_dataframe(df).filter(col("age").gt(col("score"))).collect_streaming()
        with pytest.raises(OracleError, match="Logical plan error: Column not found: 'nope'"):
            LazyFrame.from_dataframe(df).filter(both.or_(col("nope").eq(lit(1)))).collect()
    finally:
        set_extensions(False)


# ---------------------------------------------------------------- logical_plan/plan.rs + builder.rs: plan construction, schema(), validate()
def test_logical_plan_schema_and_validate():
    df = df_name_age_score()
    src = LazyFrame.from_dataframe(df)
    assert src.schema() == [("name", "String"), ("age", "Int64"), ("score", "Float64")]                 # plan.rs:446-461
    src.validate()                                                                                       # plan.rs:599-610
    assert src.select([col("name"), col("age")]).schema() == [("name", "String"), ("age", "Int64")]      # plan.rs:464-489
    assert src.select([col("name"), col("age").alias("user_age")]).schema() == [("name", "String"), ("user_age", "Int64")]   # :492-516, builder.rs:537-549
    # arithmetic: Int64 * Int64 = Int64, Float64 + Float64 = Float64, Int64 + Float64 = Float64; the name is the left operand's (plan.rs:519-548)
    arith = src.select([col("age").mul(lit(2)), col("score").add(lit(10.0)), col("age").add(col("score"))])
    assert arith.schema() == [("age", "Int64"), ("score", "Float64"), ("age", "Float64")]
    assert src.select([col("age").add(col("score")).alias("age_plus_score")]).schema() == [("age_plus_score", "Float64")]    # builder.rs:552-566
    assert src.filter(col("age").gt(lit(25))).schema() == src.schema()                                   # plan.rs:551-572
    assert src.limit(5).schema() == src.schema()                                                         # plan.rs:575-596
    src.select([col("name"), col("age")]).validate()                                                     # plan.rs:613-634
    with pytest.raises(OracleError, match="Column not found: 'invalid_column'"):                         # plan.rs:637-666
        src.select([col("name"), col("invalid_column")]).validate()
    src.filter(col("age").gt(lit(25))).validate()                                                        # plan.rs:669-684
    with pytest.raises(OracleError, match="Column not found: 'invalid'"):                                # plan.rs:687-711
        src.filter(col("invalid").gt(lit(25))).validate()
    chained = src.filter(col("age").gt(lit(25))).select([col("name"), col("score").alias("final_score")]).limit(10)   # plan.rs:715-753
    chained.validate()
    assert chained.schema() == [("name", "String"), ("final_score", "Float64")]
    nested = src.select([col("name"), col("age").mul(lit(2)).alias("double_age"), col("score")]) \
                .select([col("name"), col("double_age").add(lit(10)).alias("adjusted_age")])              # plan.rs:756-798
    nested.validate()
    assert nested.schema() == [("name", "String"), ("adjusted_age", "Int64")]


def test_lazyframe_builder_structure():
    df = df_name_age_score()
    lf = LazyFrame.from_dataframe(df)
    assert lf.describe() == "DataFrameSource"                                                            # builder.rs:165-186
    assert lf.select([col("name")]).describe() == 'Select { input: DataFrameSource, expressions: [Column("name")] }'   # :189-203
    d = lf.select([col("name"), col("age"), col("score")]).describe()                                    # :206-220
    assert d.count("Column(") == 3
    d = lf.select([col("name"), col("age").alias("user_age")]).describe()                                # :223-238
    assert 'Alias(Column("age"), "user_age")' in d
    d = lf.select([col("name"), col("age").mul(lit(2)).alias("double_age")]).describe()                  # :241-265: alias wraps a BinaryExpr
    assert 'Alias(BinaryExpr { left: Column("age"), op: Multiply, right: Literal(Int64(2)) }, "double_age")' in d
    d = lf.filter(col("age").gt(lit(30))).describe()                                                     # :270-283
    assert d == 'Filter { input: DataFrameSource, predicate: BinaryExpr { left: Column("age"), op: Gt, right: Literal(Int64(30)) } }'
    d = lf.filter(col("age").gt(lit(25)).and_(col("score").lt(lit(90.0)))).describe()                    # :286-303: the top operator is And
    assert "predicate: BinaryExpr { left: BinaryExpr {" in d and "}, op: And, right: BinaryExpr {" in d
    d = lf.filter(col("name").eq(lit("Alice"))).describe()                                               # :306-327
    assert 'BinaryExpr { left: Column("name"), op: Eq, right: Literal(String("Alice")) }' in d
    assert lf.limit(5).describe() == "Limit { input: DataFrameSource, n: 5 }"                            # :331-341
    assert lf.limit(0).describe() == "Limit { input: DataFrameSource, n: 0 }"                            # :344-354
    # select then filter / filter then select / select -> filter -> limit: the nesting follows the call order (:359-431)
    assert lf.select([col("name"), col("age")]).filter(col("age").gt(lit(25))).describe().startswith("Filter { input: Select { input: DataFrameSource")
    assert lf.filter(col("age").gt(lit(25))).select([col("name")]).describe().startswith("Select { input: Filter { input: DataFrameSource")
    d = lf.select([col("name"), col("age"), col("score")]).filter(col("age").gt(lit(25))).limit(10).describe()
    assert d.startswith("Limit { input: Filter { input: Select { input: DataFrameSource") and d.endswith(", n: 10 }")
    # a LazyFrame is a value: building on it leaves it untouched (builder.rs:619-633 clone, :636-644 debug)
    base = lf.select([col("name")])
    _ = base.limit(1)
    assert base.describe() == 'Select { input: DataFrameSource, expressions: [Column("name")] }'


# ---------------------------------------------------------------- execution/file_stream.rs
DT_I, DT_F, DT_S, DT_B, DT_N = 0, 1, 2, 3, 4      # datatypes DataType order (series.rs:126-133)


def _csv(tmp, text, name="t.csv", binary=False):
    import os
    p = os.path.join(str(tmp), name)
    with open(p, "wb") as f:
        f.write(text if binary else text.encode())
    return p


def _tmpdir():
    import tempfile
    return tempfile.mkdtemp(prefix="rvl_csv_")


def csv_test_file(tmp):  # file_stream.rs:379-388
    return _csv(tmp, "id,name,score,active\n1,Alice,85.5,true\n2,Bob,92.0,false\n3,Charlie,78.5,true\n4,,90.0,false\n5,Eve,null,true\n")


CSV_TEST_FIELDS = [("id", EX_INT64, False), ("name", EX_STRING, True), ("score", EX_FLOAT64, True), ("active", EX_BOOLEAN, False)]  # :390-397


def test_csv_file_stream_basic_and_nulls():  # file_stream.rs:399-416 (basic), :432-446 (nulls: the reference stops at num_rows)
    tmp = _tmpdir()
    sp = StreamingPhysicalPlan.csv_file_source(csv_test_file(tmp), CSV_TEST_FIELDS, 10)
    batches = sp.collect_batches()
    assert len(batches) == 1
    b = batches[0]
    assert (b.num_rows(), b.num_columns()) == (5, 4) and b.column_names() == ["id", "name", "score", "active"]
    assert b.column(0).to_list() == [1, 2, 3, 4, 5] and b.column(0).null_count == 0
    assert b.column(1).to_list() == ["Alice", "Bob", "Charlie", None, "Eve"]          # row 3: null name
    assert b.column(3).to_list() == [True, False, True, False, True]
    assert b.column(2).to_list() == [85.5, 92.0, 78.5, 90.0, None]                    # row 4: null score (corrected validity, the default)
    # batch boundaries: 2 + 2 + 1 rows
    sizes = [x.num_rows() for x in StreamingPhysicalPlan.csv_file_source(csv_test_file(tmp), CSV_TEST_FIELDS, 2).collect_batches()]
    assert sizes == [2, 2, 1]


def test_csv_adaptive_batch_size():  # file_stream.rs:418-430 (+ :346-369)
    assert calculate_adaptive_batch_size([EX_INT64, EX_STRING, EX_FLOAT64, EX_BOOLEAN]) == 100_000   # 8 MiB / 49 B, clamped
    assert calculate_adaptive_batch_size([]) == 10_000 and calculate_adaptive_batch_size([EX_NULL]) == 10_000
    assert calculate_adaptive_batch_size([EX_STRING] * 300) == 1_000                   # 9600 B per row: 873 rows, clamped up
    assert calculate_adaptive_batch_size([EX_FLOAT64] * 20) == 8 * 1024 * 1024 // 160  # 52 428, inside the clamp


def test_csv_empty_file():  # file_stream.rs:448-461: header only => no batch
    tmp = _tmpdir()
    p = _csv(tmp, "id,name\n")
    fields = [("id", EX_INT64, False), ("name", EX_STRING, True)]
    assert StreamingPhysicalPlan.csv_file_source(p, fields, 10).collect_batches() == []
    r = StreamingPhysicalPlan.csv_file_source(p, fields, 10).collect()
    assert (r.num_rows(), r.column_names()) == (0, ["id", "name"])                     # RecordBatch::empty(schema), streaming.rs:347-349
    assert StreamingPhysicalPlan.csv_file_source(_csv(tmp, "", "zero.csv"), fields, 10).collect_batches() == []   # :137-140


def test_csv_main_demo_query():  # main.rs:233-256: from_csv(';').select(3 of 4).limit(3).collect_streaming()
    tmp = _tmpdir()
    p = _csv(tmp, "Username; Identifier;First name;Last name\nbooker12;9012;Rachel;Booker\ngrey07;2070;Laura;Grey\n"
                  "johnson81;4081;Craig;Johnson\njenkins46;9346;Mary;Jenkins\nsmith79;5079;Jamie;Smith\n", "username.csv")
    schema = [("Username", DT_S), ("Identifier", DT_I), ("First_name", DT_S), ("Last_name", DT_S)]
    r = LazyFrame.from_csv(p, schema, 1000, ";").select([col("Username"), col("First_name"), col("Last_name")]).limit(3).collect_streaming()
    assert r.column_names() == ["Username", "First_name", "Last_name"] and r.num_rows() == 3
    assert r.column(0).to_list() == ["booker12", "grey07", "johnson81"] and r.column(2).to_list() == ["Booker", "Grey", "Johnson"]
    lf = LazyFrame.from_csv(p, schema, 1000, ";")
    assert lf.schema() == [("Username", "String"), ("Identifier", "Int64"), ("First_name", "String"), ("Last_name", "String")]   # plan.rs:66
    lf.validate()                                                                                                            # plan.rs:128
    with pytest.raises(OracleError, match="Logical plan error: Column not found: 'nope'"):
        lf.select([col("nope")]).collect_streaming()
    # planner.rs:45-49: the eager planner has no CSV source
    with pytest.raises(OracleError, match="Execution error: Conversion failed: CSV file source not supported in non-streaming physical planner. "
                                          "Use streaming planner instead."):
        lf.select([col("Username")]).collect()
    # streaming.rs:102-103: the file is opened when the plan executes
    with pytest.raises(OracleError, match=r"Execution error: Invalid operation: Failed to open file: No such file or directory \(os error 2\)"):
        LazyFrame.from_csv(p + ".missing", schema).limit(1).collect_streaming()


def test_csv_parse_rules():  # file_stream.rs:42-121, :153-175 — code reading: Rust str::trim / parse::<i64> / parse::<f64> / to_lowercase
    tmp = _tmpdir()
    text = ("i,f,s,b\n"
            " 7 , 1.5 ,  padded  , TRUE \n"            # fields are trimmed (:43)
            "\n   \n"                                   # blank lines are skipped and do not count towards the batch (:168-170)
            "-0,+.5e1,null,f\r\n"                       # CRLF (:162-167); '+', '.5', exponent; "null" string is NULL
            "+9223372036854775807,-inf, x　,T\n"  # i64::MAX with '+'; -inf; Unicode white space trims (NBSP, ideographic space)
            ",nan,,\n"                                  # empty = NULL for every type
            "-9223372036854775808,1e400,Null,0\n"       # i64::MIN; overflow -> inf; "Null" (capital) is a string, not NULL
            "12,1.,a b,1")                              # last line without newline; "1." is a float
    p = _csv(tmp, text)
    schema = [("i", DT_I), ("f", DT_F), ("s", DT_S), ("b", DT_B)]
    r = LazyFrame.from_csv(p, schema, 3).collect_streaming()
    assert r.num_rows() == 6
    assert r.column(0).to_list() == [7, 0, 9223372036854775807, None, -9223372036854775808, 12]
    f = r.column(1).to_list()
    assert f[0] == 1.5 and f[1] == 5.0 and f[2] == -math.inf and math.isnan(f[3]) and f[4] == math.inf and f[5] == 1.0
    assert r.column(2).to_list() == ["padded", None, "x", None, "Null", "a b"]
    assert r.column(3).to_list() == [True, False, True, None, False, True]
    assert [r.column(i).null_count for i in range(4)] == [1, 0, 2, 1]
    # a header is always consumed, whatever it holds; a Null-typed column parses nothing (:111)
    p2 = _csv(tmp, "1,2\n3,4\n", "nohdr.csv")
    r2 = LazyFrame.from_csv(p2, [("a", DT_I), ("z", DT_N)]).collect_streaming()
    assert r2.num_rows() == 1 and r2.column(0).to_list() == [3] and r2.column(1).to_list() == [None]


def test_csv_errors():  # file_stream.rs:45-52, :63-69, :79-85, :100-105, :172-176 — message text and 1-based line numbers
    tmp = _tmpdir()
    schema = [("i", DT_I), ("f", DT_F), ("b", DT_B)]

    def run(text, **kw):
        return LazyFrame.from_csv(_csv(tmp, text, binary=isinstance(text, bytes)), schema, kw.get("batch", 2)).collect_streaming()
    pre = "Execution error: Stream error: Stream execution error: Parse error: "
    with pytest.raises(OracleError, match=pre + "Line 3: Expected 3 fields, found 2"):
        run("i,f,b\n1,1.0,t\n2,2.0\n")
    with pytest.raises(OracleError, match=pre + "Line 4, field 0: Cannot parse '1_000' as Int64"):
        run("i,f,b\n1,1.0,t\n\n1_000,2.0,f\n")                      # the blank line counts as a line
    for bad in ("9223372036854775808", "+-1", "-", "1.0", "0x10", "１２"):
        with pytest.raises(OracleError, match="field 0: Cannot parse"):
            run(f"i,f,b\n{bad},1.0,t\n")
    for bad in ("1e", ".", "e5", "1.5f", "0x1p3", "in", "infinit", "1,5", "--1", "1 000"):
        if "," in bad:
            continue
        with pytest.raises(OracleError, match="field 1: Cannot parse '" + bad.replace(".", r"\.") + "' as Float64"):
            run(f"i,f,b\n1,{bad},t\n")
    with pytest.raises(OracleError, match=pre + "Line 2, field 2: Cannot parse 'yes' as Boolean"):
        run("i,f,b\n1,1.0,yes\n")
    with pytest.raises(OracleError, match="Execution error: Stream error: Stream execution error: Failed to read line 3: stream did not contain valid UTF-8"):
        run(b"i,f,b\n1,1.0,t\n2,\xff,f\n")
    # LimitStream pulls nothing once the limit is met (streaming.rs:269-271): a bad line in a batch that is never read does not surface
    ok = LazyFrame.from_csv(_csv(tmp, "i,f,b\n1,1.0,t\n2,2.0,f\n3,oops,t\n"), schema, 2).limit(2).collect_streaming()
    assert ok.column(0).to_list() == [1, 2]
    with pytest.raises(OracleError, match=pre + "Line 4, field 1: Cannot parse 'oops' as Float64"):
        LazyFrame.from_csv(_csv(tmp, "i,f,b\n1,1.0,t\n2,2.0,f\n3,oops,t\n"), schema, 2).limit(3).collect_streaming()
    with pytest.raises(OracleError, match=pre + "Line 4, field 1"):   # same file, one batch of 3: the bad line is in the first batch
        LazyFrame.from_csv(_csv(tmp, "i,f,b\n1,1.0,t\n2,2.0,f\n3,oops,t\n"), schema, 3).limit(2).collect_streaming()


def test_csv_filter_select_limit_and_validity_modes():
    tmp = _tmpdir()
    rows = [(i, None if i % 7 == 3 else i * 0.5, None if i % 5 == 0 else f"s{i}", None if i % 11 == 0 else (i % 3 == 0)) for i in range(1, 41)]
    text = "id,x,s,flag\n" + "".join(f"{i},{'' if x is None else x},{'null' if s is None else s},{'' if b is None else ('true' if b else 'F')}\n"
                                     for i, x, s, b in rows)
    p = _csv(tmp, text)
    schema = [("id", DT_I), ("x", DT_F), ("s", DT_S), ("flag", DT_B)]
    for batch in (1, 7, 16, 1000, None):
        r = LazyFrame.from_csv(p, schema, batch).filter(col("flag")).select([col("s"), col("x"), col("id")]).collect_streaming()
        keep = [t for t in rows if t[3] is True]                                    # null mask rows are dropped (take of set bits)
        assert r.column(2).to_list() == [t[0] for t in keep]
        assert r.column(1).to_list() == [t[1] for t in keep] and r.column(0).to_list() == [t[2] for t in keep]
        r = LazyFrame.from_csv(p, schema, batch).filter(col("flag")).select([col("id")]).limit(5).collect_streaming()
        assert r.column(0).to_list() == [t[0] for t in keep][:5]
        r = LazyFrame.from_csv(p, schema, batch).limit(9).collect_streaming()
        assert r.column(0).to_list() == list(range(1, 10)) and r.column(3).to_list() == [t[3] for t in rows[:9]]
    # the reference's inverted validity for Int64 / Float64 columns holding a null (file_stream.rs:233-239 vs primitive.rs:31-33)
    set_csv_reference_validity(True)
    try:
        r = LazyFrame.from_csv(p, schema, 1000).collect_streaming()
        assert r.column(1).to_list() == [0.0 if t[1] is None else None for t in rows]     # only the NULL fields come out valid (as 0.0)
        assert r.column(0).to_list() == [t[0] for t in rows]                              # no null in the column: no bitmap, untouched
        assert r.column(2).to_list() == [t[2] for t in rows] and r.column(3).to_list() == [t[3] for t in rows]   # String / Boolean are right
        small = LazyFrame.from_csv(p, schema, 2).collect_streaming()                      # per batch: batches without a null stay valid
        exp = []
        for k in range(0, 40, 2):
            pair = rows[k:k + 2]
            anyn = any(t[1] is None for t in pair)
            exp += [(0.0 if t[1] is None else None) if anyn else t[1] for t in pair]
        assert small.column(1).to_list() == exp
    finally:
        set_csv_reference_validity(False)


# ---------------------------------------------------------------- inner join: physical_plan/plan.rs:174-284
def users_orders():  # main.rs:120-165
    users = DataFrame.new([("user_id", [1, 2, 3, 4]), ("name", ["Alice", "Bob", "Charlie", "Diana"]), ("city", ["Rome", "Milan", "Naples", "Turin"])])
    orders = DataFrame.new([("order_id", [101, 102, 103, 104, 105]), ("user_id", [1, 2, 1, 3, 2]), ("amount", [29.99, 15.5, 45.0, 8.75, 12.99])])
    return users, orders


def test_join_main_demo_queries():  # main.rs:170-196
    users, orders = users_orders()
    lf = LazyFrame.from_dataframe(users).inner_join(LazyFrame.from_dataframe(orders), "user_id", "user_id")
    r = lf.collect()
    # the probe (right) side's columns come first, then the build (left) side's without its key (plan.rs:218-250); rows in probe order
    assert r.column_names() == ["order_id", "user_id", "amount", "name", "city"]
    assert r.to_dict() == {"order_id": [101, 102, 103, 104, 105], "user_id": [1, 2, 1, 3, 2], "amount": [29.99, 15.5, 45.0, 8.75, 12.99],
                           "name": ["Alice", "Bob", "Alice", "Charlie", "Bob"], "city": ["Rome", "Milan", "Rome", "Naples", "Milan"]}
    r = lf.select([col("name"), col("amount"), col("city")]).collect()                       # main.rs:186-196
    assert r.to_dict() == {"name": ["Alice", "Bob", "Alice", "Charlie", "Bob"], "amount": [29.99, 15.5, 45.0, 8.75, 12.99],
                           "city": ["Rome", "Milan", "Rome", "Naples", "Milan"]}
    # logical schema (logical_plan/plan.rs:79-111): left columns, then the right ones without the right key
    assert lf.schema() == [("user_id", "Int64"), ("name", "String"), ("city", "String"), ("order_id", "Int64"), ("amount", "Float64")]
    assert lf.plan_shape() == "Join(Source, Source)"
    assert lf.describe() == 'Join { left: DataFrameSource, right: DataFrameSource, left_key: "user_id", right_key: "user_id", join_type: Inner }'
    # the optimizer recurses into both sides (optimizer.rs:50-61)
    lf2 = LazyFrame.from_dataframe(users).filter(col("user_id").gt(lit(1))).select([col("user_id"), col("name")]) \
        .inner_join(LazyFrame.from_dataframe(orders).filter(col("amount").gt(lit(10.0))), "user_id", "user_id")
    assert lf2.plan_shape() == "Join(Filter(Select(Source)), Filter(Source))"
    assert lf2.collect().to_dict() == {"order_id": [102, 105], "user_id": [2, 2], "amount": [15.5, 12.99], "name": ["Bob", "Bob"]}


def test_join_plan_errors():  # logical_plan/plan.rs:156-200, physical_plan/streaming.rs:128-131
    users, orders = users_orders()
    lu, lo = LazyFrame.from_dataframe(users), LazyFrame.from_dataframe(orders)
    with pytest.raises(OracleError, match="Logical plan error: Column not found: 'nope'"):
        lu.inner_join(lo, "nope", "user_id").collect()
    with pytest.raises(OracleError, match="Logical plan error: Column not found: 'missing'"):
        lu.inner_join(lo, "user_id", "missing").collect()
    with pytest.raises(OracleError, match="Logical plan error: Incompatible join key types: 'String' and 'Int64'"):
        lu.inner_join(lo, "name", "user_id").collect()
    lu.inner_join(lo, "user_id", "amount").validate()                      # Int64 with Float64 is "comparable" (series.rs:150-151) ...
    assert lu.inner_join(lo, "user_id", "amount").collect().height() == 0  # ... but an Int64 never equals a Float64 (series.rs:87-98)
    with pytest.raises(OracleError, match="not yet implemented: Streaming hash join not yet implemented"):
        lu.inner_join(lo, "user_id", "user_id").collect_streaming()
    # a key column on both sides with the same non-key name: "_right" goes on the BUILD (left) column (plan.rs:237-241)
    a = DataFrame.new([("k", [1, 2]), ("v", [10, 20])])
    b = DataFrame.new([("k", [2, 1, 2]), ("v", [7.5, 8.5, 9.5])])
    r = LazyFrame.from_dataframe(a).inner_join(LazyFrame.from_dataframe(b), "k", "k").collect()
    assert r.column_names() == ["k", "v", "v_right"] and r.to_dict() == {"k": [2, 1, 2], "v": [7.5, 8.5, 9.5], "v_right": [20, 10, 20]}
    # different key names: the probe keeps its key, the build key is dropped; a clash with the generated name is a DataFrame error
    c = DataFrame.new([("id", [1, 2]), ("v", [1, 2]), ("v_right", [3, 4])])
    with pytest.raises(OracleError, match="Execution error: DataFrame error: Duplicate column name: 'v_right'"):
        LazyFrame.from_dataframe(a).inner_join(LazyFrame.from_dataframe(c), "k", "id").collect()


def test_join_key_semantics():  # HashMap<AnyValue, Vec<usize>> (plan.rs:186-205) with AnyValue's Hash / Eq (series.rs:73-98) — code reading
    nan = float("nan")
    left = DataFrame.new([("k", [1, None, 2, 1, None]), ("l", ["a", "b", "c", "d", "e"])])
    right = DataFrame.new([("k", [None, 1, 3, 2]), ("r", [10.0, 20.0, 30.0, 40.0])])
    r = LazyFrame.from_dataframe(left).inner_join(LazyFrame.from_dataframe(right), "k", "k").collect()
    # Null == Null (series.rs:89); every probe row meets its build rows in ascending build order
    assert r.to_dict() == {"k": [None, None, 1, 1, 2], "r": [10.0, 10.0, 20.0, 20.0, 40.0], "l": ["b", "e", "a", "d", "c"]}
    # Float64 keys: bit-equal values meet, NaN meets nothing; an all-null result column collapses to dtype Null (Series::new)
    fl = DataFrame.new([("x", [1.5, nan, 2.5, None]), ("tag", [None, None, "t", None])])
    fr = DataFrame.new([("x", [nan, 1.5, None, 1.5])])
    r = LazyFrame.from_dataframe(fl).inner_join(LazyFrame.from_dataframe(fr), "x", "x").collect()
    assert r.height() == 3 and r.column("tag") == [None, None, None] and r.dtypes() == ["Float64", "Null"]
    assert [v for v in r.column("x")] == [1.5, None, 1.5]
    # String and Boolean keys
    sl = DataFrame.new([("s", ["x", "yy", "x", "", None]), ("n", [1, 2, 3, 4, 5])])
    sr = DataFrame.new([("s", ["yy", "", "zzz", "x"]), ("b", [True, False, True, None])])
    r = LazyFrame.from_dataframe(sl).inner_join(LazyFrame.from_dataframe(sr), "s", "s").collect()
    assert r.to_dict() == {"s": ["yy", "", "x", "x"], "b": [True, False, None, None], "n": [2, 4, 1, 3]}
    r = LazyFrame.from_dataframe(sr).inner_join(LazyFrame.from_dataframe(sr), "b", "b").select([col("s"), col("s_right")]).collect()
    assert r.to_dict() == {"s": ["yy", "yy", "", "zzz", "zzz", "x"], "s_right": ["yy", "zzz", "", "yy", "zzz", "x"]}
    # no pair at all: empty series keep their dtypes (create_empty_join_result, plan.rs:256-283)
    r = LazyFrame.from_dataframe(left).inner_join(LazyFrame.from_dataframe(DataFrame.new([("k", [7, 8]), ("z", ["p", "q"])])), "k", "k").collect()
    assert (r.height(), r.column_names(), r.dtypes()) == (0, ["k", "z", "l"], ["Int64", "String", "String"])
    # a Float64 Series holding Int64 values (series.rs:210-212): every row keeps its own type as a key
    ml = DataFrame.new([("m", [1.0, 1, 2.0, 2]), ("i", [0, 1, 2, 3])])
    mr = DataFrame.new([("m", [1, 2.0, 3])])
    r = LazyFrame.from_dataframe(ml).inner_join(LazyFrame.from_dataframe(mr), "m", "m").collect()
    assert r.to_dict()["i"] == [1, 2]


# ---------------------------------------------------------------- RecordBatch::validate / memory_size / new_unchecked, RecordBatchBuilder
def _id_name_active_columns():  # record_batch.rs:585-604 (create_test_schema / create_test_columns)
    return [Array.from_list([1, 2, 3], EX_INT64), Array.from_list(["Alice", None, "Charlie"], EX_STRING), Array.from_list([True, False, True], EX_BOOLEAN)]


ID_NAME_ACTIVE = (["id", "name", "active"], [EX_INT64, EX_STRING, EX_BOOLEAN])


def test_rb_memory_size_and_validate():  # record_batch.rs:952-981
    rb = rb_id_name_active()
    assert rb.memory_size() > 0                                                      # :952-959
    assert rb.memory_size() == 24 + 24 + 3 * 16 + 3 * 8 + 3 * 20 + 1                 # :380-400: schema, Vec, 3 ArrayRefs, i64 / String estimate / packed bits
    rb.validate()                                                                    # :962-968
    cols = _id_name_active_columns()
    bad = RecordBatch.new_unchecked(ID_NAME_ACTIVE[0], [cols[0], Array.from_list(["Alice"], EX_STRING), cols[2]], 3, ID_NAME_ACTIVE[1])   # :971-981
    with pytest.raises(OracleError, match="Column 1 has length 1 but expected 3"):
        bad.validate()
    with pytest.raises(OracleError, match="Schema has 3 fields but 2 columns present"):
        RecordBatch.new_unchecked(ID_NAME_ACTIVE[0], cols[:2], 3, ID_NAME_ACTIVE[1]).validate()
    with pytest.raises(OracleError, match="Column 1 has type Boolean but schema expects String"):
        RecordBatch.new_unchecked(ID_NAME_ACTIVE[0], [cols[0], cols[2], cols[2]], 3, ID_NAME_ACTIVE[1]).validate()
    ok = RecordBatch.new_unchecked(ID_NAME_ACTIVE[0], cols, 3, ID_NAME_ACTIVE[1])
    ok.validate()
    assert ok.column_by_name("name").to_list() == ["Alice", None, "Charlie"] and ok.column_by_name("nope") is None   # :673-688


def test_rb_builder():  # record_batch.rs:984-1046
    cols = _id_name_active_columns()
    b = RecordBatchBuilder(*ID_NAME_ACTIVE)                                          # :984-1013
    assert b.num_columns() == 0 and not b.is_complete()
    for c in cols:
        b.add_column(c)
    assert b.num_columns() == 3 and b.is_complete()
    rb = b.finish()
    assert (rb.num_rows(), rb.num_columns()) == (3, 3) and rb.column(1).to_list() == ["Alice", None, "Charlie"]
    with pytest.raises(OracleError, match="Cannot add more columns than schema defines"):   # :519-521
        b.add_column(cols[0])
    assert RecordBatchBuilder.with_capacity(ID_NAME_ACTIVE[0], ID_NAME_ACTIVE[1], 1000).num_columns() == 0   # :1016-1021
    b = RecordBatchBuilder(*ID_NAME_ACTIVE)                                          # :1024-1032: wrong type for the first field
    with pytest.raises(OracleError, match="Column type String doesn't match expected type Int64"):
        b.add_column(Array.from_list(["test"], EX_STRING))
    b.add_column(cols[0])
    with pytest.raises(OracleError, match="Column length 1 doesn't match expected length 3"):   # :532-541
        b.add_column(Array.from_list(["only one"], EX_STRING))
    with pytest.raises(OracleError, match="Expected 3 columns but only 1 provided"):           # :1035-1046
        b.finish()


def test_datatype_predicates():  # series.rs:368-395
    assert dtype_is_numeric(DT_I) and dtype_is_numeric(DT_F) and not dtype_is_numeric(DT_S) and not dtype_is_numeric(DT_B) and not dtype_is_numeric(DT_N)
    assert dtype_is_comparable_with(DT_I, DT_I) and dtype_is_comparable_with(DT_S, DT_S)          # same types
    assert dtype_is_comparable_with(DT_I, DT_F) and dtype_is_comparable_with(DT_F, DT_I)          # numeric with numeric
    assert dtype_is_comparable_with(DT_N, DT_I) and dtype_is_comparable_with(DT_S, DT_N)          # Null with everything
    assert not dtype_is_comparable_with(DT_S, DT_B) and not dtype_is_comparable_with(DT_B, DT_I)  # different non-numeric types
