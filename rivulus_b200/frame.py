"""rivulus_b200.frame — the reference's user-facing API (LazyFrame / DataFrame / Expr / RecordBatch /
StreamingPhysicalPlan) on the GPU.

A thin ctypes front-end to the C++ host layer (rivulus_b200/host/, librivulus_host.so), which mirrors
/root/reference/src (logical_plan/builder.rs, datatypes/*, expressions/expr.rs, execution/record_batch.rs,
physical_plan/streaming.rs) and executes every query through the C ABI of include/rivulus_gpu.h.

    from rivulus_b200.frame import DataFrame, LazyFrame, col, lit
    df = DataFrame.new([("name", ["Alice", "Bob", "Charlie"]), ("age", [25, 30, 35])])
    out = LazyFrame.from_dataframe(df).filter(col("age").gt(lit(25))).select([col("name")]).collect()
    out.to_dict()   # {"name": ["Bob", "Charlie"]}

There is no CPU execution path: collect() / collect_streaming() need the built libraries and a B200.
Errors carry the reference's Display text (QueryError / ExecutionError / StreamError ...); `.panic` marks what the
reference reports by panicking.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "librivulus_host.so")

# AnyValue tags and dtype enums (same numbering as the reference's declaration order)
NULL, INT64, FLOAT64, STRING, BOOLEAN = 0, 1, 2, 3, 4
DT_INT64, DT_FLOAT64, DT_STRING, DT_BOOLEAN, DT_NULL = 0, 1, 2, 3, 4          # datatypes/series.rs:126-133
DT_NAMES = ["Int64", "Float64", "String", "Boolean", "Null"]
EX_NULL, EX_BOOLEAN, EX_INT64, EX_FLOAT64, EX_STRING = 0, 1, 2, 3, 4          # execution/schema.rs:1-8
OPS = {"+": 0, "-": 1, "*": 2, "/": 3, "==": 4, "!=": 5, "<": 6, ">": 7, "<=": 8, ">=": 9, "and": 10, "or": 11}  # expr.rs:15-29

# every symbol the host library exports for this front-end (tests check they resolve)
HOST_SYMBOLS = [
    "rvh_last_error", "rvh_launch_count", "rvh_set_stream_fusion", "rvh_set_extensions", "rvh_dtype_is_numeric", "rvh_dtype_is_comparable_with",
    "rvh_dfb_new", "rvh_dfb_add_series", "rvh_dfb_add_empty_series", "rvh_dfb_add_i64", "rvh_dfb_add_f64", "rvh_dfb_add_bool_bits",
    "rvh_dfb_add_strings", "rvh_dfb_finish", "rvh_synth_df", "rvh_df_free", "rvh_df_width", "rvh_df_height", "rvh_df_col_name", "rvh_df_col_dtype",
    "rvh_df_col_len", "rvh_df_col_str_bytes", "rvh_df_col_export", "rvh_df_col_buffers",
    "rvh_expr_col", "rvh_expr_lit", "rvh_expr_binary", "rvh_expr_alias", "rvh_expr_free",
    "rvh_lf_from_df", "rvh_lf_from_csv", "rvh_set_csv_reference_validity", "rvh_set_csv_threads", "rvh_csv_adaptive_batch_size", "rvh_sp_csv_source", "rvh_csv_parse_dump", "rvh_free", "rvh_lf_inner_join", "rvh_lf_select", "rvh_lf_filter", "rvh_lf_limit", "rvh_lf_free", "rvh_lf_collect", "rvh_lf_collect_streaming",
    "rvh_rb_new_unchecked", "rvh_rb_validate", "rvh_rb_memory_size", "rvh_rbb_new", "rvh_rbb_add_column", "rvh_rbb_finish", "rvh_rbb_num_columns",
    "rvh_rbb_is_complete", "rvh_rbb_free",
    "rvh_lf_plan_shape", "rvh_lf_schema", "rvh_lf_validate", "rvh_lf_describe",
    "rvh_rb_try_new", "rvh_rb_free", "rvh_rb_num_rows", "rvh_rb_num_columns", "rvh_rb_col_name", "rvh_rb_col_dtype", "rvh_rb_col_export",
    "rvh_rb_slice", "rvh_rb_take", "rvh_rb_select", "rvh_rb_select_by_name", "rvh_rb_filter", "rvh_rb_concat", "rvh_rb_empty_like",
    "rvh_sp_memory_source", "rvh_sp_dataframe_source", "rvh_sp_filter", "rvh_sp_select", "rvh_sp_limit", "rvh_sp_free", "rvh_sp_collect",
    "rvh_sp_collect_batches", "rvh_rbv_len", "rvh_rbv_get", "rvh_rbv_free",
]


class ColExport(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("length", C.c_int64), ("offset", C.c_int64), ("null_count", C.c_int64),
                ("values", C.c_void_p), ("values_len", C.c_int64),
                ("validity", C.c_void_p), ("validity_len", C.c_int64),
                ("offsets", C.c_void_p), ("offsets_len", C.c_int64),
                ("data", C.c_void_p), ("data_len", C.c_int64)]


_lib = None


def lib():
    """Load librivulus_host.so (which links librivulus_gpu.so).  Raises ImportError if not built — never falls back."""
    global _lib
    if _lib is None:
        capi.lib()  # the kernel library first: a clear error if it is missing
        if not os.path.exists(HOST_LIB_PATH):
            raise ImportError(f"{HOST_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(or `make -C rivulus_b200/host`).  rivulus_b200 has no CPU fallback.")
        L = C.CDLL(HOST_LIB_PATH)
        L.rvh_last_error.restype = C.c_char_p
        for name in ("rvh_dfb_new", "rvh_expr_col", "rvh_expr_lit", "rvh_expr_binary", "rvh_expr_alias", "rvh_lf_from_df", "rvh_lf_from_csv", "rvh_sp_csv_source", "rvh_lf_inner_join", "rvh_rbb_new", "rvh_lf_select",
                     "rvh_lf_filter", "rvh_lf_limit", "rvh_sp_memory_source", "rvh_sp_dataframe_source", "rvh_sp_filter", "rvh_sp_select",
                     "rvh_sp_limit", "rvh_rbv_get"):
            getattr(L, name).restype = C.c_void_p
        for name in ("rvh_df_col_name", "rvh_rb_col_name"):
            getattr(L, name).restype = C.c_char_p
        for name in ("rvh_df_height", "rvh_df_col_len", "rvh_df_col_str_bytes", "rvh_rb_num_rows", "rvh_launch_count", "rvh_csv_adaptive_batch_size", "rvh_rb_memory_size"):
            getattr(L, name).restype = C.c_int64
        _lib = L
    return _lib


class RivulusError(Exception):
    """A reference `Err(..)`; str(e) is the Display text.  `.panic` marks a Rust panic."""

    def __init__(self, msg, panic=False):
        super().__init__(msg)
        self.panic = panic


def _check(rc):
    if rc != 0:
        msg = lib().rvh_last_error().decode(errors="replace")
        raise RivulusError(msg, panic=(rc == 2))


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def _vp(h):
    return C.c_void_p(h)


def launch_count(device: int = 0) -> int:
    """Kernels launched so far by the default context of `device`."""
    return lib().rvh_launch_count(device)


def set_stream_fusion(on: bool) -> None:
    """collect_streaming(): run [Limit] -> [Select] -> [Filter] over a DataFrame as one rvl_stream pipeline (default) or, when off,
    as one operator object per plan node like the reference's pull chain — same result either way."""
    lib().rvh_set_stream_fusion(1 if on else 0)


def set_extensions(on: bool) -> None:
    """Opt-in extension (off = the reference's behaviour and error text): And / Or over comparison leaves in collect() and
    collect_streaming(), comparison predicates in collect_streaming() (rivulus.hpp: set_extensions)."""
    lib().rvh_set_extensions(1 if on else 0)


def dtype_is_numeric(d: int) -> bool:                      # series.rs:136-142 over DT_*
    return bool(lib().rvh_dtype_is_numeric(d))


def dtype_is_comparable_with(a: int, b: int) -> bool:      # series.rs:144-159
    return bool(lib().rvh_dtype_is_comparable_with(a, b))


def set_csv_reference_validity(on: bool) -> None:
    """True = reproduce the reference's inverted validity of Int64 / Float64 CSV columns that hold a null (file_stream.rs:213-240);
    default False = null fields are null."""
    lib().rvh_set_csv_reference_validity(1 if on else 0)


def set_csv_threads(n: int) -> None:
    """Parse threads per CSV reader: -1 (default) = up to 8 for files of at least 8 MiB, none below; 0 = the calling thread parses."""
    lib().rvh_set_csv_threads(int(n))


def calculate_adaptive_batch_size(exec_dtypes) -> int:
    """file_stream.rs:346-369 over execution dtypes (EX_*)."""
    a = (C.c_int * max(len(exec_dtypes), 1))(*exec_dtypes)
    return int(lib().rvh_csv_adaptive_batch_size(len(exec_dtypes), a))


def csv_parse_dump(path, exec_dtypes, batch_size=None, delimiter=None, reference_validity=False) -> str:
    """CPU-only probe of the CSV parser (tests): every batch rendered as text, see rvh_csv_parse_dump in host_capi.cpp."""
    a = (C.c_int * max(len(exec_dtypes), 1))(*exec_dtypes)
    out = C.c_void_p()
    rc = lib().rvh_csv_parse_dump(str(path).encode(), len(exec_dtypes), a, C.c_int64(-1 if batch_size is None else batch_size),
                                  None if delimiter is None else delimiter.encode(), 1 if reference_validity else 0, C.byref(out))
    if rc != 0:
        raise RivulusError(lib().rvh_last_error().decode(errors="replace"))
    try:
        return C.string_at(out.value).decode()
    finally:
        lib().rvh_free(out)


def any_tag(v):
    if v is None:
        return NULL
    if isinstance(v, (bool, np.bool_)):
        return BOOLEAN
    if isinstance(v, (int, np.integer)):
        return INT64
    if isinstance(v, (float, np.floating)):
        return FLOAT64
    if isinstance(v, (str, bytes)):
        return STRING
    raise TypeError(type(v))


def _lit_args(v):
    t = any_tag(v)
    s = v.encode() if isinstance(v, str) else (v if isinstance(v, bytes) else b"")
    return (t, int(v) if t == INT64 else 0, float(v) if t == FLOAT64 else 0.0, s, len(s), int(bool(v)) if t == BOOLEAN else 0)


# ------------------------------------------------------------------------------------------ DataFrame
class DataFrame:
    """datatypes/dataframe.rs.  Columns are held in Arrow layout on the host (rivulus.hpp: Series)."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rvh_df_free(_vp(self._h))
            self._h = None

    @staticmethod
    def new(series: Sequence[tuple]) -> "DataFrame":
        """DataFrame::new(vec![Series::new(name, values)...]); items: (name, [AnyValue-like scalars, None = Null]) or
        (name, [], dtype) for Series::empty."""
        L = lib()
        b = L.rvh_dfb_new()
        for item in series:
            name, vals = item[0], item[1]
            if len(vals) == 0 and len(item) > 2:
                _check(L.rvh_dfb_add_empty_series(_vp(b), name.encode(), item[2]))
                continue
            n = len(vals)
            tags = np.array([any_tag(v) for v in vals], dtype=np.uint8)
            i64 = np.array([int(v) if t == INT64 else 0 for v, t in zip(vals, tags)], dtype=np.int64)
            f64 = np.array([float(v) if t == FLOAT64 else 0.0 for v, t in zip(vals, tags)], dtype=np.float64)
            b8 = np.array([1 if (t == BOOLEAN and v) else 0 for v, t in zip(vals, tags)], dtype=np.uint8)
            enc = [(v.encode() if isinstance(v, str) else v) if t == STRING else b"" for v, t in zip(vals, tags)]
            off = np.zeros(n + 1, dtype=np.int32)
            if n:
                off[1:] = np.cumsum([len(e) for e in enc])
            data = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8)
            rc = L.rvh_dfb_add_series(_vp(b), name.encode(), C.c_int64(n), _ptr(tags, C.c_uint8), _ptr(i64, C.c_int64),
                                      _ptr(f64, C.c_double), _ptr(b8, C.c_uint8), _ptr(off, C.c_int32), _ptr(data, C.c_uint8))
            if rc != 0:
                msg = L.rvh_last_error().decode(errors="replace")
                out = C.c_void_p()
                L.rvh_dfb_finish(_vp(b), C.byref(out))
                if out.value:
                    L.rvh_df_free(out)
                raise RivulusError(msg, panic=(rc == 2))
        out = C.c_void_p()
        _check(L.rvh_dfb_finish(_vp(b), C.byref(out)))
        return DataFrame(out.value)

    @staticmethod
    def from_columns(columns: Sequence[tuple]) -> "DataFrame":
        """Columnar ingestion (no per-value objects): items are (name, capi.Column) with host numpy buffers at offset 0."""
        L = lib()
        b = L.rvh_dfb_new()
        for name, c in columns:
            n = c.length
            vb = _ptr(c.validity, C.c_uint8) if c.validity is not None else None
            if c.dtype == capi.INT64:
                v = np.ascontiguousarray(c.values, dtype=np.int64)
                rc = L.rvh_dfb_add_i64(_vp(b), name.encode(), C.c_int64(n), _ptr(v, C.c_int64), vb)
            elif c.dtype == capi.FLOAT64:
                v = np.ascontiguousarray(c.values, dtype=np.float64)
                rc = L.rvh_dfb_add_f64(_vp(b), name.encode(), C.c_int64(n), _ptr(v, C.c_double), vb)
            elif c.dtype == capi.BOOLEAN:
                v = np.ascontiguousarray(c.values, dtype=np.uint8)
                rc = L.rvh_dfb_add_bool_bits(_vp(b), name.encode(), C.c_int64(n), _ptr(v, C.c_uint8), vb)
            elif c.dtype == capi.STRING:
                o = np.ascontiguousarray(c.offsets, dtype=np.int32)
                d = np.ascontiguousarray(c.data if c.data is not None and c.data.size else np.zeros(1, np.uint8), dtype=np.uint8)
                rc = L.rvh_dfb_add_strings(_vp(b), name.encode(), C.c_int64(n), _ptr(o, C.c_int32), _ptr(d, C.c_uint8), vb)
            else:
                raise TypeError("unsupported column dtype")
            _check(rc)
        out = C.c_void_p()
        _check(L.rvh_dfb_finish(_vp(b), C.byref(out)))
        return DataFrame(out.value)

    @staticmethod
    def synth(cols: Sequence[tuple], n: int, row0: int = 0) -> "DataFrame":
        """cols: [(name, kind, col_id, null_pct)] from include/rivulus_synth.h's counter-based generator (BASELINE configs)."""
        L = lib()
        names = (C.c_char_p * len(cols))(*[c[0].encode() for c in cols])
        kinds = (C.c_int * len(cols))(*[c[1] for c in cols])
        ids = (C.c_uint32 * len(cols))(*[c[2] for c in cols])
        nulls = (C.c_uint32 * len(cols))(*[c[3] for c in cols])
        out = C.c_void_p()
        _check(L.rvh_synth_df(len(cols), names, kinds, ids, nulls, C.c_uint64(row0), C.c_int64(n), C.byref(out)))
        return DataFrame(out.value)

    def width(self):
        return lib().rvh_df_width(_vp(self._h))

    def height(self):
        return lib().rvh_df_height(_vp(self._h))

    def shape(self):
        return (self.height(), self.width())

    def column_names(self):
        return [lib().rvh_df_col_name(_vp(self._h), i).decode() for i in range(self.width())]

    def dtypes(self):
        return [DT_NAMES[lib().rvh_df_col_dtype(_vp(self._h), i)] for i in range(self.width())]

    def column_raw(self, i):
        """(tags, i64, f64, b8, str_off, str_data) numpy arrays for column i."""
        L = lib()
        n = L.rvh_df_col_len(_vp(self._h), i)
        sb = L.rvh_df_col_str_bytes(_vp(self._h), i)
        tags = np.zeros(n, np.uint8); i64 = np.zeros(n, np.int64); f64 = np.zeros(n, np.float64); b8 = np.zeros(n, np.uint8)
        off = np.zeros(n + 1, np.int32); data = np.zeros(max(sb, 1), np.uint8)
        L.rvh_df_col_export(_vp(self._h), i, _ptr(tags, C.c_uint8), _ptr(i64, C.c_int64), _ptr(f64, C.c_double),
                            _ptr(b8, C.c_uint8), _ptr(off, C.c_int32), _ptr(data, C.c_uint8))
        return tags, i64, f64, b8, off, data[:sb]

    def column_buffers(self, i) -> "ArrowColumn":
        """Raw Arrow buffers of column i (copied out)."""
        e = ColExport()
        lib().rvh_df_col_buffers(_vp(self._h), i, C.byref(e))
        return _export(e)

    def column(self, key) -> list:
        """Column as a python list of scalars (None = Null)."""
        i = self.column_names().index(key) if isinstance(key, str) else key
        tags, i64, f64, b8, off, data = self.column_raw(i)
        out = []
        raw = data.tobytes()
        for r, t in enumerate(tags):
            if t == NULL: out.append(None)
            elif t == INT64: out.append(int(i64[r]))
            elif t == FLOAT64: out.append(float(f64[r]))
            elif t == BOOLEAN: out.append(bool(b8[r]))
            else: out.append(raw[off[r]:off[r + 1]].decode())
        return out

    def to_dict(self):
        return {n: self.column(i) for i, n in enumerate(self.column_names())}


# ------------------------------------------------------------------------------------------ Expr / LazyFrame
class Expr:
    """expressions/expr.rs"""

    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rvh_expr_free(_vp(self._h)); self._h = None

    def alias(self, name):
        return Expr(lib().rvh_expr_alias(_vp(self._h), name.encode()))

    def _bin(self, op, other):
        return Expr(lib().rvh_expr_binary(_vp(self._h), OPS[op], _vp(other._h)))

    def add(self, o): return self._bin("+", o)
    def sub(self, o): return self._bin("-", o)
    def mul(self, o): return self._bin("*", o)
    def div(self, o): return self._bin("/", o)
    def eq(self, o): return self._bin("==", o)
    def neq(self, o): return self._bin("!=", o)
    def lt(self, o): return self._bin("<", o)
    def gt(self, o): return self._bin(">", o)
    def lte(self, o): return self._bin("<=", o)
    def gte(self, o): return self._bin(">=", o)
    def and_(self, o): return self._bin("and", o)
    def or_(self, o): return self._bin("or", o)


def col(name) -> Expr:
    return Expr(lib().rvh_expr_col(name.encode()))


def lit(v) -> Expr:
    t, i, f, s, sl, b = _lit_args(v)
    return Expr(lib().rvh_expr_lit(t, C.c_int64(i), C.c_double(f), s, C.c_int64(sl), b))


class LazyFrame:
    """logical_plan/builder.rs:11-114"""

    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rvh_lf_free(_vp(self._h)); self._h = None

    @staticmethod
    def from_dataframe(df: DataFrame):
        return LazyFrame(lib().rvh_lf_from_df(_vp(df._h)))

    @staticmethod
    def from_csv(path, schema, batch_size=None, delimiter=None):
        """builder.rs:41-55.  schema: [(name, DT_*)]; delimiter: one character."""
        names = (C.c_char_p * max(len(schema), 1))(*[n.encode() for n, _ in schema])
        dts = (C.c_int * max(len(schema), 1))(*[d for _, d in schema])
        return LazyFrame(lib().rvh_lf_from_csv(str(path).encode(), len(schema), names, dts, C.c_int64(-1 if batch_size is None else batch_size),
                                               None if delimiter is None else delimiter.encode()))

    def inner_join(self, right: "LazyFrame", left_key: str, right_key: str):
        """builder.rs:84-94"""
        return LazyFrame(lib().rvh_lf_inner_join(_vp(self._h), _vp(right._h), left_key.encode(), right_key.encode()))

    def select(self, exprs: List[Expr]):
        arr = (C.c_void_p * max(len(exprs), 1))(*[e._h for e in exprs])
        return LazyFrame(lib().rvh_lf_select(_vp(self._h), len(exprs), arr))

    def filter(self, pred: Expr):
        return LazyFrame(lib().rvh_lf_filter(_vp(self._h), _vp(pred._h)))

    def limit(self, n: int):
        return LazyFrame(lib().rvh_lf_limit(_vp(self._h), C.c_int64(n)))

    def collect(self) -> DataFrame:
        """Eager engine semantics (physical_plan/plan.rs), executed by the fused GPU operator."""
        out = C.c_void_p()
        _check(lib().rvh_lf_collect(_vp(self._h), C.byref(out)))
        return DataFrame(out.value)

    def collect_streaming(self) -> "RecordBatch":
        """Streaming engine semantics (physical_plan/streaming.rs), executed on the GPU."""
        out = C.c_void_p()
        _check(lib().rvh_lf_collect_streaming(_vp(self._h), C.byref(out)))
        return RecordBatch(out.value)

    def plan_shape(self) -> str:
        buf = C.create_string_buffer(256)
        _check(lib().rvh_lf_plan_shape(_vp(self._h), buf, 256))
        return buf.value.decode()

    def schema(self):
        """logical_plan.schema() of the plan as built: [(name, dtype name)] (logical_plan/plan.rs:63-113)."""
        buf = C.create_string_buffer(4096)
        _check(lib().rvh_lf_schema(_vp(self._h), buf, 4096))
        return [tuple(x.split(":")) for x in buf.value.decode().split(",") if x]

    def validate(self):
        """logical_plan.validate() of the plan as built (logical_plan/plan.rs:115-202); raises on ColumnNotFound."""
        _check(lib().rvh_lf_validate(_vp(self._h)))

    def describe(self) -> str:
        """Debug-style dump of the plan as built (node kinds, expression trees)."""
        buf = C.create_string_buffer(8192)
        _check(lib().rvh_lf_describe(_vp(self._h), buf, 8192))
        return buf.value.decode()


# ------------------------------------------------------------------------------------------ Arrow-layout columns
@dataclass
class ArrowColumn:
    """Raw buffers of one column, exactly as the reference would hold them (host copy)."""
    dtype: int                      # execution::schema::DataType order
    length: int
    offset: int
    null_count: int
    values: Optional[np.ndarray]
    validity: Optional[np.ndarray]  # None when the bitmap is absent
    offsets: Optional[np.ndarray]
    data: Optional[np.ndarray]

    def to_list(self):
        def bit(buf, i):
            return (int(buf[i >> 3]) >> (i & 7)) & 1
        out = []
        for r in range(self.length):
            li = self.offset + r
            if self.dtype == EX_NULL or (self.validity is not None and not bit(self.validity, li)):
                out.append(None)
            elif self.dtype == EX_INT64: out.append(int(self.values[li]))
            elif self.dtype == EX_FLOAT64: out.append(float(self.values[li]))
            elif self.dtype == EX_BOOLEAN: out.append(bool(bit(self.values, li)))
            else: out.append(self.data[self.offsets[li]:self.offsets[li + 1]].tobytes().decode())
        return out


def _np_from(ptr, n, dtype):
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def _export(e: ColExport) -> ArrowColumn:
    vals = None
    if e.dtype == EX_INT64: vals = _np_from(e.values, e.values_len, np.int64)
    elif e.dtype == EX_FLOAT64: vals = _np_from(e.values, e.values_len, np.float64)
    elif e.dtype == EX_BOOLEAN: vals = _np_from(e.values, e.values_len, np.uint8)
    return ArrowColumn(e.dtype, e.length, e.offset, e.null_count, vals,
                       _np_from(e.validity, e.validity_len, np.uint8) if e.validity else None,
                       _np_from(e.offsets, e.offsets_len, np.int32) if e.dtype == EX_STRING else None,
                       _np_from(e.data, e.data_len, np.uint8) if e.dtype == EX_STRING else None)


class RecordBatch:
    """execution/record_batch.rs — device resident; column() copies one column back to the host."""

    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rvh_rb_free(_vp(self._h)); self._h = None

    @staticmethod
    def try_new(names: Sequence[str], columns: Sequence[capi.Column], schema_dtypes: Optional[Sequence[int]] = None,
                schema_names: Optional[Sequence[str]] = None) -> "RecordBatch":
        """RecordBatch::try_new(schema, arrays) over host buffers (uploaded).  schema_dtypes / schema_names let a test
        provoke the reference's mismatch errors."""
        sn = list(schema_names if schema_names is not None else names)
        sd = list(schema_dtypes if schema_dtypes is not None else [c.dtype for c in columns])
        nm = (C.c_char_p * max(len(sn), 1))(*[n.encode() for n in sn])
        dt = (C.c_int * max(len(sd), 1))(*sd)
        arr = (capi.RvlColumn * max(len(columns), 1))(*[c.as_struct() for c in columns])
        out = C.c_void_p()
        _check(lib().rvh_rb_try_new(len(sn), nm, dt, len(columns), arr, C.byref(out)))
        return RecordBatch(out.value)

    @staticmethod
    def new_unchecked(names, columns, num_rows, schema_dtypes):
        """RecordBatch::new_unchecked (record_batch.rs:60-66): no checks; validate() reports what is wrong."""
        nm = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
        dt = (C.c_int * max(len(schema_dtypes), 1))(*schema_dtypes)
        arr = (capi.RvlColumn * max(len(columns), 1))(*[c.as_struct() for c in columns])
        out = C.c_void_p()
        _check(lib().rvh_rb_new_unchecked(len(names), nm, dt, len(columns), arr, C.c_int64(num_rows), C.byref(out)))
        return RecordBatch(out.value)

    def validate(self): _check(lib().rvh_rb_validate(_vp(self._h)))            # record_batch.rs:348-378 (raises with the Err text)
    def memory_size(self): return int(lib().rvh_rb_memory_size(_vp(self._h)))   # :380-400

    def column_by_name(self, name):                                             # :84-86
        names = self.column_names()
        return self.column(names.index(name)) if name in names else None

    def num_rows(self): return lib().rvh_rb_num_rows(_vp(self._h))
    def num_columns(self): return lib().rvh_rb_num_columns(_vp(self._h))
    def column_names(self): return [lib().rvh_rb_col_name(_vp(self._h), i).decode() for i in range(self.num_columns())]
    def column_dtypes(self): return [lib().rvh_rb_col_dtype(_vp(self._h), i) for i in range(self.num_columns())]

    def column(self, i) -> ArrowColumn:
        if isinstance(i, str):
            i = self.column_names().index(i)
        e = ColExport()
        _check(lib().rvh_rb_col_export(_vp(self._h), i, C.byref(e)))
        return _export(e)

    def columns(self): return [self.column(i) for i in range(self.num_columns())]
    def to_dict(self): return {n: self.column(i).to_list() for i, n in enumerate(self.column_names())}

    def _op(self, fn, *args):
        out = C.c_void_p()
        _check(fn(_vp(self._h), *args, C.byref(out)))
        return RecordBatch(out.value)

    def slice(self, off, length): return self._op(lib().rvh_rb_slice, C.c_int64(off), C.c_int64(length))

    def take(self, idx):
        """RecordBatch::take (record_batch.rs:108-129)."""
        a = np.ascontiguousarray(idx, dtype=np.int64)
        return self._op(lib().rvh_rb_take, _ptr(a, C.c_int64) if a.size else None, C.c_int64(a.size))

    def filter(self, predicate_batch: "RecordBatch", predicate_column: int = 0):
        """filter(&self, predicate): the predicate array is column `predicate_column` of `predicate_batch`."""
        return self._op(lib().rvh_rb_filter, _vp(predicate_batch._h), predicate_column)

    def select_columns(self, idx):
        a = (C.c_int32 * max(len(idx), 1))(*idx)
        return self._op(lib().rvh_rb_select, a, len(idx))

    def select_columns_by_name(self, names):
        a = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
        return self._op(lib().rvh_rb_select_by_name, a, len(names))

    @staticmethod
    def concat(batches):
        out = C.c_void_p()
        arr = (C.c_void_p * max(len(batches), 1))(*[b._h for b in batches])
        _check(lib().rvh_rb_concat(arr, len(batches), C.byref(out)))
        return RecordBatch(out.value)

    def empty_like(self): return self._op(lib().rvh_rb_empty_like)


class RecordBatchBuilder:
    """execution/record_batch.rs:495-573 over host columns (uploaded by finish())."""

    def __init__(self, names, schema_dtypes, capacity=0):
        nm = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
        dt = (C.c_int * max(len(schema_dtypes), 1))(*schema_dtypes)
        self._h = lib().rvh_rbb_new(len(names), nm, dt)
        self._keep = []          # the builder borrows the columns' host buffers until finish()

    @staticmethod
    def with_capacity(names, schema_dtypes, capacity): return RecordBatchBuilder(names, schema_dtypes, capacity)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rvh_rbb_free(_vp(self._h)); self._h = None

    def add_column(self, column: capi.Column):
        st = column.as_struct()
        _check(lib().rvh_rbb_add_column(_vp(self._h), C.byref(st)))
        self._keep.append((column, st))

    def finish(self) -> "RecordBatch":
        out = C.c_void_p()
        _check(lib().rvh_rbb_finish(_vp(self._h), C.byref(out)))
        return RecordBatch(out.value)

    def num_columns(self): return lib().rvh_rbb_num_columns(_vp(self._h))
    def is_complete(self): return bool(lib().rvh_rbb_is_complete(_vp(self._h)))


class StreamingPhysicalPlan:
    """physical_plan/streaming.rs:28-133, 235-243, 290-333"""

    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().rvh_sp_free(_vp(self._h)); self._h = None

    @staticmethod
    def memory_source(batches: Sequence[RecordBatch]):
        arr = (C.c_void_p * max(len(batches), 1))(*[b._h for b in batches])
        return StreamingPhysicalPlan(lib().rvh_sp_memory_source(arr, len(batches)))

    @staticmethod
    def dataframe_source(df: DataFrame, batch_size: int):
        h = lib().rvh_sp_dataframe_source(_vp(df._h), C.c_int64(batch_size))
        if not h:
            raise RivulusError(lib().rvh_last_error().decode(errors="replace"))
        return StreamingPhysicalPlan(h)

    @staticmethod
    def csv_file_source(path, fields, batch_size=None, delimiter=None):
        """streaming.rs:299-311.  fields: [(name, EX_*, nullable)]."""
        names = (C.c_char_p * max(len(fields), 1))(*[f[0].encode() for f in fields])
        dts = (C.c_int * max(len(fields), 1))(*[f[1] for f in fields])
        nul = (C.c_int * max(len(fields), 1))(*[1 if f[2] else 0 for f in fields])
        return StreamingPhysicalPlan(lib().rvh_sp_csv_source(str(path).encode(), len(fields), names, dts, nul,
                                                             C.c_int64(-1 if batch_size is None else batch_size),
                                                             None if delimiter is None else delimiter.encode()))

    def filter(self, column: str):
        return StreamingPhysicalPlan(lib().rvh_sp_filter(_vp(self._h), column.encode()))

    def select(self, names: Sequence[str]):
        a = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
        return StreamingPhysicalPlan(lib().rvh_sp_select(_vp(self._h), a, len(names)))

    def limit(self, n: int):
        return StreamingPhysicalPlan(lib().rvh_sp_limit(_vp(self._h), C.c_int64(n)))

    def collect(self) -> RecordBatch:
        out = C.c_void_p()
        _check(lib().rvh_sp_collect(_vp(self._h), C.byref(out)))
        return RecordBatch(out.value)

    def collect_batches(self) -> List[RecordBatch]:
        out = C.c_void_p()
        _check(lib().rvh_sp_collect_batches(_vp(self._h), C.byref(out)))
        L = lib()
        n = L.rvh_rbv_len(out)
        res = [RecordBatch(L.rvh_rbv_get(out, i)) for i in range(n)]
        L.rvh_rbv_free(out)
        return res
