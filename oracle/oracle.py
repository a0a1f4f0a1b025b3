"""ctypes front-end to the CPU oracle (oracle/lib/liboracle.so) — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (rivulus_b200/) never does.

The classes mirror the reference's public names (AnyValue/Series/DataFrame/Expr/LazyFrame,
RecordBatch, StreamingPhysicalPlan) so the golden tests in tests/ read like the reference's own
unit tests (file:line cited per test).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "liboracle.so")

# AnyValue tags (oracle AnyValue::Tag) and dtype enums
NULL, INT64, FLOAT64, STRING, BOOLEAN = 0, 1, 2, 3, 4
# datatypes::series::DataType order (series.rs:126-133)
DT_INT64, DT_FLOAT64, DT_STRING, DT_BOOLEAN, DT_NULL = 0, 1, 2, 3, 4
DT_NAMES = ["Int64", "Float64", "String", "Boolean", "Null"]
# execution::schema::DataType order (schema.rs:1-8)
EX_NULL, EX_BOOLEAN, EX_INT64, EX_FLOAT64, EX_STRING = 0, 1, 2, 3, 4
# expr.rs:15-29
OPS = {"+": 0, "-": 1, "*": 2, "/": 3, "==": 4, "!=": 5, "<": 6, ">": 7, "<=": 8, ">=": 9, "and": 10, "or": 11}


def build(force: bool = False) -> str:
    """Compile the oracle with g++ (idempotent)."""
    srcs = [os.path.join(_HERE, f) for f in ("rivulus_oracle.cpp", "oracle_capi.cpp", "rivulus_oracle.hpp")]
    srcs.append(os.path.join(_HERE, "..", "include", "rivulus_synth.h"))
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs if os.path.exists(s))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class ColExport(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("length", C.c_int64), ("offset", C.c_int64), ("null_count", C.c_int64),
                ("values", C.c_void_p), ("values_len", C.c_int64),
                ("validity", C.c_void_p), ("validity_len", C.c_int64),
                ("offsets", C.c_void_p), ("offsets_len", C.c_int64),
                ("data", C.c_void_p), ("data_len", C.c_int64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_dfb_new", "orc_expr_col", "orc_expr_lit", "orc_expr_binary", "orc_expr_alias", "orc_lf_from_df", "orc_lf_from_csv", "orc_sp_csv_source", "orc_lf_inner_join", "orc_rbb_new",
                     "orc_lf_select", "orc_lf_filter", "orc_lf_limit", "orc_arr_i64", "orc_arr_f64", "orc_arr_bool",
                     "orc_arr_str", "orc_arr_null", "orc_arr_i64_new", "orc_arr_f64_new", "orc_arr_bool_new",
                     "orc_sp_memory_source", "orc_sp_dataframe_source", "orc_sp_filter", "orc_sp_select", "orc_sp_limit",
                     "orc_rbv_get"):
            getattr(L, name).restype = C.c_void_p
        for name in ("orc_df_col_name", "orc_rb_col_name"):
            getattr(L, name).restype = C.c_char_p
        for name in ("orc_df_height", "orc_df_col_len", "orc_df_col_str_bytes", "orc_arr_len", "orc_arr_null_count",
                     "orc_rb_num_rows", "orc_csv_adaptive_batch_size", "orc_rb_memory_size"):
            getattr(L, name).restype = C.c_int64
        L.orc_time_eager_filter_select.restype = C.c_double
        L.orc_time_eager_filter_select_mt.restype = C.c_double
        _lib = L
    return _lib


class OracleError(Exception):
    """A reference `Err(..)`; str(e) is the Display text.  `.panic` marks a Rust panic."""

    def __init__(self, msg, panic=False):
        super().__init__(msg)
        self.panic = panic


def _check(rc):
    if rc != 0:
        raise OracleError(lib().orc_last_error().decode(), panic=(rc == 2))


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def _vp(h):
    return C.c_void_p(h)


# ------------------------------------------------------------------------------------------ AnyValue helpers
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
    """datatypes/dataframe.rs — built from python lists of AnyValue-like scalars (None = Null)."""

    def __init__(self, handle):
        self._h = handle

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_df_free(_vp(self._h))
            self._h = None

    @staticmethod
    def new(series: Sequence[tuple]) -> "DataFrame":
        """series: [(name, [values...])] or (name, [], dtype) for Series::empty."""
        L = lib()
        b = L.orc_dfb_new()
        for item in series:
            name, vals = item[0], item[1]
            if len(vals) == 0 and len(item) > 2:
                _check(L.orc_dfb_add_empty_series(_vp(b), name.encode(), item[2]))
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
            _check(L.orc_dfb_add_series(_vp(b), name.encode(), C.c_int64(n), _ptr(tags, C.c_uint8), _ptr(i64, C.c_int64),
                                        _ptr(f64, C.c_double), _ptr(b8, C.c_uint8), _ptr(off, C.c_int32), _ptr(data, C.c_uint8)))
        out = C.c_void_p()
        _check(L.orc_dfb_finish(_vp(b), C.byref(out)))
        return DataFrame(out.value)

    @staticmethod
    def synth(cols: Sequence[tuple], n: int, row0: int = 0) -> "DataFrame":
        """cols: [(name, kind, col_id, null_pct)] using include/rivulus_synth.h's generator."""
        L = lib()
        names = (C.c_char_p * len(cols))(*[c[0].encode() for c in cols])
        kinds = (C.c_int * len(cols))(*[c[1] for c in cols])
        ids = (C.c_uint32 * len(cols))(*[c[2] for c in cols])
        nulls = (C.c_uint32 * len(cols))(*[c[3] for c in cols])
        out = C.c_void_p()
        _check(L.orc_synth_df(len(cols), names, kinds, ids, nulls, C.c_uint64(row0), C.c_int64(n), C.byref(out)))
        return DataFrame(out.value)

    def width(self):
        return lib().orc_df_width(_vp(self._h))

    def height(self):
        return lib().orc_df_height(_vp(self._h))

    def column_names(self):
        return [lib().orc_df_col_name(_vp(self._h), i).decode() for i in range(self.width())]

    def dtypes(self):
        return [DT_NAMES[lib().orc_df_col_dtype(_vp(self._h), i)] for i in range(self.width())]

    def column_raw(self, i):
        """(tags, i64, f64, b8, str_off, str_data) numpy arrays for column i."""
        L = lib()
        n = L.orc_df_col_len(_vp(self._h), i)
        sb = L.orc_df_col_str_bytes(_vp(self._h), i)
        tags = np.zeros(n, np.uint8); i64 = np.zeros(n, np.int64); f64 = np.zeros(n, np.float64); b8 = np.zeros(n, np.uint8)
        off = np.zeros(n + 1, np.int32); data = np.zeros(max(sb, 1), np.uint8)
        L.orc_df_col_export(_vp(self._h), i, _ptr(tags, C.c_uint8), _ptr(i64, C.c_int64), _ptr(f64, C.c_double),
                            _ptr(b8, C.c_uint8), _ptr(off, C.c_int32), _ptr(data, C.c_uint8))
        return tags, i64, f64, b8, off, data[:sb]

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
    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_expr_free(_vp(self._h)); self._h = None

    def alias(self, name):
        return Expr(lib().orc_expr_alias(_vp(self._h), name.encode()))

    def _bin(self, op, other):
        return Expr(lib().orc_expr_binary(_vp(self._h), OPS[op], _vp(other._h)))

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


def set_extensions(on: bool) -> None:
    """Opt-in extension (off = reference behaviour): And / Or over comparison leaves in collect() and collect_streaming(),
    comparison predicates in collect_streaming().  See rivulus_oracle.hpp for the definition the oracle restates."""
    lib().orc_set_extensions(1 if on else 0)


def set_csv_reference_validity(on: bool) -> None:
    """True = the reference's inverted validity of Int64 / Float64 CSV columns holding a null (file_stream.rs:213-240); default False."""
    lib().orc_set_csv_reference_validity(1 if on else 0)


def calculate_adaptive_batch_size(exec_dtypes) -> int:
    a = (C.c_int * max(len(exec_dtypes), 1))(*exec_dtypes)
    return int(lib().orc_csv_adaptive_batch_size(len(exec_dtypes), a))


def dtype_is_numeric(d: int) -> bool:                      # series.rs:136-142 over DT_*
    return bool(lib().orc_dtype_is_numeric(d))


def dtype_is_comparable_with(a: int, b: int) -> bool:      # series.rs:144-159
    return bool(lib().orc_dtype_is_comparable_with(a, b))


def col(name) -> Expr:
    return Expr(lib().orc_expr_col(name.encode()))


def lit(v) -> Expr:
    t, i, f, s, sl, b = _lit_args(v)
    return Expr(lib().orc_expr_lit(t, C.c_int64(i), C.c_double(f), s, C.c_int64(sl), b))


class LazyFrame:
    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_lf_free(_vp(self._h)); self._h = None

    @staticmethod
    def from_dataframe(df: DataFrame):
        return LazyFrame(lib().orc_lf_from_df(_vp(df._h)))

    @staticmethod
    def from_csv(path, schema, batch_size=None, delimiter=None):
        names = (C.c_char_p * max(len(schema), 1))(*[n.encode() for n, _ in schema])
        dts = (C.c_int * max(len(schema), 1))(*[d for _, d in schema])
        return LazyFrame(lib().orc_lf_from_csv(str(path).encode(), len(schema), names, dts, C.c_int64(-1 if batch_size is None else batch_size),
                                               None if delimiter is None else delimiter.encode()))

    def select(self, exprs: List[Expr]):
        arr = (C.c_void_p * len(exprs))(*[e._h for e in exprs])
        return LazyFrame(lib().orc_lf_select(_vp(self._h), len(exprs), arr))

    def filter(self, pred: Expr):
        return LazyFrame(lib().orc_lf_filter(_vp(self._h), _vp(pred._h)))

    def inner_join(self, right: "LazyFrame", left_key: str, right_key: str):
        return LazyFrame(lib().orc_lf_inner_join(_vp(self._h), _vp(right._h), left_key.encode(), right_key.encode()))

    def limit(self, n: int):
        return LazyFrame(lib().orc_lf_limit(_vp(self._h), C.c_int64(n)))

    def collect(self) -> DataFrame:
        out = C.c_void_p()
        _check(lib().orc_lf_collect(_vp(self._h), C.byref(out)))
        return DataFrame(out.value)

    def collect_streaming(self) -> "RecordBatch":
        out = C.c_void_p()
        _check(lib().orc_lf_collect_streaming(_vp(self._h), C.byref(out)))
        return RecordBatch(out.value)

    def plan_shape(self) -> str:
        buf = C.create_string_buffer(256)
        lib().orc_lf_plan_shape(_vp(self._h), buf, 256)
        return buf.value.decode()

    def schema(self):
        """logical_plan.schema() of the plan as built: [(name, dtype name)] (logical_plan/plan.rs:63-113)."""
        buf = C.create_string_buffer(4096)
        _check(lib().orc_lf_schema(_vp(self._h), buf, 4096))
        return [tuple(x.split(":")) for x in buf.value.decode().split(",") if x]

    def validate(self):
        """logical_plan.validate() of the plan as built (logical_plan/plan.rs:115-202); raises on ColumnNotFound."""
        _check(lib().orc_lf_validate(_vp(self._h)))

    def describe(self) -> str:
        """Debug-style dump of the plan as built (node kinds, expression trees)."""
        buf = C.create_string_buffer(8192)
        _check(lib().orc_lf_describe(_vp(self._h), buf, 8192))
        return buf.value.decode()


# ------------------------------------------------------------------------------------------ Arrow-layout arrays
@dataclass
class ArrowColumn:
    """Raw buffers of one column, exactly as the reference would hold them."""
    dtype: int                      # execution::schema::DataType order
    length: int
    offset: int
    null_count: int
    values: Optional[np.ndarray]    # int64/float64 values (whole buffer) or uint8 value-bitmap bytes
    validity: Optional[np.ndarray]  # uint8 bitmap bytes (whole buffer) or None when the bitmap is absent
    offsets: Optional[np.ndarray]   # int32
    data: Optional[np.ndarray]      # uint8

    def to_list(self):
        """Logical values (None = null), honouring offset/validity."""
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


def pack_bits(bools) -> np.ndarray:
    """LSB-first packing (bitmap.rs:44-59)."""
    b = np.asarray(bools, dtype=np.uint8)
    return np.packbits(b, bitorder="little") if b.size else np.zeros(0, np.uint8)


class Array:
    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_arr_free(_vp(self._h)); self._h = None

    # --- raw-buffer constructors (whole buffers + (offset, length) window) ---
    @staticmethod
    def i64(values, validity_bits=None, offset=0, length=None):
        v = np.ascontiguousarray(values, dtype=np.int64)
        vb = None if validity_bits is None else np.ascontiguousarray(validity_bits, dtype=np.uint8)
        n = len(v) if length is None else length
        return Array(lib().orc_arr_i64(_ptr(v, C.c_int64), C.c_int64(len(v)), _ptr(vb, C.c_uint8), C.c_int64(offset), C.c_int64(n)))

    @staticmethod
    def f64(values, validity_bits=None, offset=0, length=None):
        v = np.ascontiguousarray(values, dtype=np.float64)
        vb = None if validity_bits is None else np.ascontiguousarray(validity_bits, dtype=np.uint8)
        n = len(v) if length is None else length
        return Array(lib().orc_arr_f64(_ptr(v, C.c_double), C.c_int64(len(v)), _ptr(vb, C.c_uint8), C.c_int64(offset), C.c_int64(n)))

    @staticmethod
    def boolean(value_bits, nbits, validity_bits=None, offset=0, length=None):
        v = np.ascontiguousarray(value_bits, dtype=np.uint8)
        vb = None if validity_bits is None else np.ascontiguousarray(validity_bits, dtype=np.uint8)
        n = nbits if length is None else length
        return Array(lib().orc_arr_bool(_ptr(v, C.c_uint8), C.c_int64(nbits), _ptr(vb, C.c_uint8), C.c_int64(offset), C.c_int64(n)))

    @staticmethod
    def string(offsets, data, validity_bits=None, offset=0, length=None):
        o = np.ascontiguousarray(offsets, dtype=np.int32)
        d = np.ascontiguousarray(data, dtype=np.uint8)
        if d.size == 0:
            d = np.zeros(1, np.uint8); dl = 0
        else:
            dl = len(d)
        vb = None if validity_bits is None else np.ascontiguousarray(validity_bits, dtype=np.uint8)
        ns = len(o) - 1
        n = ns if length is None else length
        return Array(lib().orc_arr_str(_ptr(o, C.c_int32), C.c_int64(ns), _ptr(d, C.c_uint8), C.c_int64(dl), _ptr(vb, C.c_uint8),
                                       C.c_int64(offset), C.c_int64(n)))

    @staticmethod
    def null(n):
        return Array(lib().orc_arr_null(C.c_int64(n)))

    # --- reference constructors over Option<T> lists (exercise the builders) ---
    @staticmethod
    def from_list(vals, dtype):
        L = lib()
        n = len(vals)
        valid = np.array([0 if v is None else 1 for v in vals], dtype=np.uint8)
        v8 = _ptr(valid, C.c_uint8) if (valid == 0).any() or dtype in (EX_BOOLEAN, EX_STRING) else None
        if dtype == EX_INT64:
            a = np.array([0 if v is None else v for v in vals], dtype=np.int64)
            return Array(L.orc_arr_i64_new(_ptr(a, C.c_int64), C.c_int64(n), v8))
        if dtype == EX_FLOAT64:
            a = np.array([0.0 if v is None else v for v in vals], dtype=np.float64)
            return Array(L.orc_arr_f64_new(_ptr(a, C.c_double), C.c_int64(n), v8))
        if dtype == EX_BOOLEAN:
            a = np.array([1 if v else 0 for v in vals], dtype=np.uint8)
            return Array(L.orc_arr_bool_new(_ptr(a, C.c_uint8), C.c_int64(n), v8))
        if dtype == EX_STRING:
            enc = [b"" if v is None else (v.encode() if isinstance(v, str) else v) for v in vals]
            off = np.zeros(n + 1, np.int32)
            if n:
                off[1:] = np.cumsum([len(e) for e in enc])
            data = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8)
            out = C.c_void_p()
            _check(L.orc_arr_str_new(_ptr(off, C.c_int32), _ptr(data, C.c_uint8), C.c_int64(n), v8, C.byref(out)))
            return Array(out.value)
        return Array.null(n)

    def slice(self, off, length):
        out = C.c_void_p()
        _check(lib().orc_arr_slice(_vp(self._h), C.c_int64(off), C.c_int64(length), C.byref(out)))
        return Array(out.value)

    def len(self):
        return lib().orc_arr_len(_vp(self._h))

    def null_count(self):
        return lib().orc_arr_null_count(_vp(self._h))

    def export(self) -> ArrowColumn:
        e = ColExport()
        lib().orc_arr_export(_vp(self._h), C.byref(e))
        return _export(e)

    def to_list(self):
        return self.export().to_list()


class RecordBatch:
    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_rb_free(_vp(self._h)); self._h = None

    @staticmethod
    def try_new(names: Sequence[str], arrays: Sequence[Array], schema_dtypes: Optional[Sequence[int]] = None,
                schema_names: Optional[Sequence[str]] = None):
        out = C.c_void_p()
        arrs = (C.c_void_p * len(arrays))(*[a._h for a in arrays])
        if schema_dtypes is None:
            nm = (C.c_char_p * len(names))(*[n.encode() for n in names])
            _check(lib().orc_rb_new(len(arrays), nm, arrs, C.byref(out)))
        else:
            sn = schema_names if schema_names is not None else names
            nm = (C.c_char_p * len(sn))(*[n.encode() for n in sn])
            dt = (C.c_int * len(schema_dtypes))(*schema_dtypes)
            _check(lib().orc_rb_try_new(len(sn), nm, dt, len(arrays), arrs, C.byref(out)))
        return RecordBatch(out.value)

    @staticmethod
    def new_unchecked(names, arrays, num_rows, schema_dtypes):
        """RecordBatch::new_unchecked (record_batch.rs:60-66): no checks; validate() reports what is wrong."""
        out = C.c_void_p()
        arrs = (C.c_void_p * max(len(arrays), 1))(*[a._h for a in arrays])
        nm = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
        dt = (C.c_int * max(len(schema_dtypes), 1))(*schema_dtypes)
        _check(lib().orc_rb_new_unchecked(len(names), nm, dt, len(arrays), arrs, C.c_int64(num_rows), C.byref(out)))
        return RecordBatch(out.value)

    def validate(self): _check(lib().orc_rb_validate(_vp(self._h)))            # record_batch.rs:348-378 (raises with the Err text)
    def memory_size(self): return int(lib().orc_rb_memory_size(_vp(self._h)))   # :380-400

    def column_by_name(self, name):                                             # :84-86
        names = self.column_names()
        return self.column(names.index(name)) if name in names else None

    def num_rows(self): return lib().orc_rb_num_rows(_vp(self._h))
    def num_columns(self): return lib().orc_rb_num_columns(_vp(self._h))
    def column_names(self): return [lib().orc_rb_col_name(_vp(self._h), i).decode() for i in range(self.num_columns())]

    def column(self, i) -> ArrowColumn:
        if isinstance(i, str):
            i = self.column_names().index(i)
        e = ColExport()
        lib().orc_rb_col_export(_vp(self._h), i, C.byref(e))
        return _export(e)

    def columns(self): return [self.column(i) for i in range(self.num_columns())]
    def to_dict(self): return {n: self.column(i).to_list() for i, n in enumerate(self.column_names())}

    def _op(self, fn, *args):
        out = C.c_void_p()
        _check(fn(_vp(self._h), *args, C.byref(out)))
        return RecordBatch(out.value)

    def slice(self, off, length): return self._op(lib().orc_rb_slice, C.c_int64(off), C.c_int64(length))

    def take(self, idx):
        a = np.ascontiguousarray(idx, dtype=np.int64)
        return self._op(lib().orc_rb_take, _ptr(a, C.c_int64) if a.size else None, C.c_int64(a.size))

    def filter(self, mask: Array): return self._op(lib().orc_rb_filter, _vp(mask._h))

    def select_columns(self, idx):
        a = (C.c_int32 * len(idx))(*idx)
        return self._op(lib().orc_rb_select, a, len(idx))

    def select_columns_by_name(self, names):
        a = (C.c_char_p * len(names))(*[n.encode() for n in names])
        return self._op(lib().orc_rb_select_by_name, a, len(names))

    @staticmethod
    def concat(batches):
        out = C.c_void_p()
        arr = (C.c_void_p * len(batches))(*[b._h for b in batches])
        _check(lib().orc_rb_concat(arr, len(batches), C.byref(out)))
        return RecordBatch(out.value)

    def empty_like(self): return self._op(lib().orc_rb_empty_like)

    def filter_project_cmp(self, pred_col: int, op: str, literal, proj: Sequence[int], limit: int = -1):
        """Fused-operator oracle: eager truth table (plan.rs:112-130) + record_batch.rs kernels."""
        t, i, f, s, sl, b = _lit_args(literal)
        p = (C.c_int32 * len(proj))(*proj)
        return self._op(lib().orc_rb_filter_project_cmp, pred_col, OPS[op], t, C.c_int64(i), C.c_double(f), s, C.c_int64(sl), b,
                        p, len(proj), C.c_int64(limit))

    def filter_project_mask(self, mask_col: int, proj: Sequence[int], limit: int = -1):
        p = (C.c_int32 * len(proj))(*proj)
        return self._op(lib().orc_rb_filter_project_mask, mask_col, p, len(proj), C.c_int64(limit))


class RecordBatchBuilder:
    """execution/record_batch.rs:495-573"""

    def __init__(self, names, schema_dtypes, capacity=0):
        nm = (C.c_char_p * max(len(names), 1))(*[n.encode() for n in names])
        dt = (C.c_int * max(len(schema_dtypes), 1))(*schema_dtypes)
        self._h = lib().orc_rbb_new(len(names), nm, dt)

    @staticmethod
    def with_capacity(names, schema_dtypes, capacity): return RecordBatchBuilder(names, schema_dtypes, capacity)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_rbb_free(_vp(self._h)); self._h = None

    def add_column(self, array): _check(lib().orc_rbb_add_column(_vp(self._h), _vp(array._h)))

    def finish(self):
        out = C.c_void_p()
        _check(lib().orc_rbb_finish(_vp(self._h), C.byref(out)))
        return RecordBatch(out.value)

    def num_columns(self): return lib().orc_rbb_num_columns(_vp(self._h))
    def is_complete(self): return bool(lib().orc_rbb_is_complete(_vp(self._h)))


class StreamingPhysicalPlan:
    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_sp_free(_vp(self._h)); self._h = None

    @staticmethod
    def memory_source(batches):
        arr = (C.c_void_p * max(len(batches), 1))(*[b._h for b in batches])
        return StreamingPhysicalPlan(lib().orc_sp_memory_source(arr, len(batches)))

    @staticmethod
    def dataframe_source(df: DataFrame, batch_size: int):
        return StreamingPhysicalPlan(lib().orc_sp_dataframe_source(_vp(df._h), C.c_int64(batch_size)))

    @staticmethod
    def csv_file_source(path, fields, batch_size=None, delimiter=None):
        names = (C.c_char_p * max(len(fields), 1))(*[f[0].encode() for f in fields])
        dts = (C.c_int * max(len(fields), 1))(*[f[1] for f in fields])
        nul = (C.c_int * max(len(fields), 1))(*[1 if f[2] else 0 for f in fields])
        return StreamingPhysicalPlan(lib().orc_sp_csv_source(str(path).encode(), len(fields), names, dts, nul,
                                                             C.c_int64(-1 if batch_size is None else batch_size),
                                                             None if delimiter is None else delimiter.encode()))

    def filter(self, colname): return StreamingPhysicalPlan(lib().orc_sp_filter(_vp(self._h), colname.encode()))

    def select(self, names):
        a = (C.c_char_p * len(names))(*[n.encode() for n in names])
        return StreamingPhysicalPlan(lib().orc_sp_select(_vp(self._h), a, len(names)))

    def limit(self, n): return StreamingPhysicalPlan(lib().orc_sp_limit(_vp(self._h), C.c_int64(n)))

    def collect(self) -> RecordBatch:
        out = C.c_void_p()
        _check(lib().orc_sp_collect(_vp(self._h), C.byref(out)))
        return RecordBatch(out.value)

    def collect_batches(self) -> List[RecordBatch]:
        out = C.c_void_p()
        _check(lib().orc_sp_collect_batches(_vp(self._h), C.byref(out)))
        L = lib()
        res = [RecordBatch(L.orc_rbv_get(_vp(out.value), i)) for i in range(L.orc_rbv_len(_vp(out.value)))]
        L.orc_rbv_free(_vp(out.value))
        return res


# ------------------------------------------------------------------------------------------ scalar probes
def _sc(v):
    t = any_tag(v)
    return (t, C.c_int64(int(v) if t == INT64 else 0), C.c_double(float(v) if t == FLOAT64 else 0.0),
            (v.encode() if isinstance(v, str) else b""), int(bool(v)) if t == BOOLEAN else 0)


def any_eq(a, b) -> bool:
    return bool(lib().orc_any_eq(*_sc(a), *_sc(b)))


def any_partial_cmp(a, b):
    """-1 / 0 / 1, or None (Rust `None`)."""
    r = lib().orc_any_partial_cmp(*_sc(a), *_sc(b))
    return None if r == 2 else r


def eval_cmp(row, op: str, literal) -> bool:
    return bool(lib().orc_eval_cmp(*_sc(row), OPS[op], *_sc(literal)))


# ------------------------------------------------------------------------------------------ timed CPU baselines
def time_eager_filter_select(df: DataFrame, pred_col: str, op: str, literal, proj: Sequence[str]):
    """Seconds for from_dataframe(df).filter(pred).select(proj).collect() on 1 thread; returns (secs, rows_out)."""
    t, i, f, s, sl, b = _lit_args(literal)
    names = (C.c_char_p * max(len(proj), 1))(*[p.encode() for p in proj])
    rows = C.c_int64()
    secs = lib().orc_time_eager_filter_select(_vp(df._h), pred_col.encode(), OPS[op], t, C.c_int64(i), C.c_double(f),
                                              names, len(proj), C.byref(rows))
    if secs < 0:
        raise OracleError(lib().orc_last_error().decode())
    return secs, rows.value


def time_eager_filter_select_mt(dfs: Sequence[DataFrame], pred_col: str, op: str, literal, proj: Sequence[str]):
    """Same query on len(dfs) shards concurrently (one reference instance per thread); (wall secs, rows_out)."""
    t, i, f, s, sl, b = _lit_args(literal)
    names = (C.c_char_p * max(len(proj), 1))(*[p.encode() for p in proj])
    hs = (C.c_void_p * len(dfs))(*[d._h for d in dfs])
    rows = C.c_int64()
    secs = lib().orc_time_eager_filter_select_mt(hs, len(dfs), pred_col.encode(), OPS[op], t, C.c_int64(i), C.c_double(f),
                                                 names, len(proj), C.byref(rows))
    if secs < 0:
        raise OracleError(lib().orc_last_error().decode())
    return secs, rows.value


def synth_filter_checksums(n: int, row0: int, pred_kind: int, pred_col_id: int, op: str, literal, proj: Sequence[tuple],
                           threads: int = 0, limit: int = -1):
    """(count, [checksum per projected (kind, col_id)]) of the filtered synthetic table, straight from the generator."""
    if threads <= 0:
        threads = os.cpu_count() or 1
    t, i, f, s, sl, b = _lit_args(literal)
    kinds = (C.c_int * max(len(proj), 1))(*[p[0] for p in proj])
    ids = (C.c_uint32 * max(len(proj), 1))(*[p[1] for p in proj])
    count = C.c_int64()
    sums = (C.c_uint64 * max(len(proj), 1))()
    _check(lib().orc_synth_filter_checksums(C.c_int64(n), C.c_uint64(row0), pred_kind, C.c_uint32(pred_col_id), OPS[op], t,
                                            C.c_int64(i), C.c_double(f), len(proj), kinds, ids, threads, C.c_int64(limit),
                                            C.byref(count), sums))
    return count.value, [sums[k] for k in range(len(proj))]


def synth_filter_checksums_nulls(n: int, row0: int, pred: tuple, op: str, literal, proj: Sequence[tuple], threads: int = 0, limit: int = -1):
    """pred = (kind, col_id, null_pct); proj = [(kind, col_id, null_pct)].  Returns (count, checksums, null_counts, string_bytes)
    of the filtered synthetic table with nulls / strings, straight from the generator (see orc_synth_filter_checksums_nulls)."""
    if threads <= 0:
        threads = os.cpu_count() or 1
    t, i, f, s, sl, b = _lit_args(literal)
    m = max(len(proj), 1)
    kinds = (C.c_int * m)(*[p[0] for p in proj])
    ids = (C.c_uint32 * m)(*[p[1] for p in proj])
    nulls = (C.c_uint32 * m)(*[p[2] for p in proj])
    count = C.c_int64()
    sums = (C.c_uint64 * m)(); nc = (C.c_int64 * m)(); nb = (C.c_int64 * m)()
    _check(lib().orc_synth_filter_checksums_nulls(C.c_int64(n), C.c_uint64(row0), pred[0], C.c_uint32(pred[1]), C.c_uint32(pred[2]), OPS[op], t,
                                                  C.c_int64(i), C.c_double(f), len(proj), kinds, ids, nulls, threads, C.c_int64(limit),
                                                  C.byref(count), sums, nc, nb))
    k = len(proj)
    return count.value, [sums[j] for j in range(k)], [nc[j] for j in range(k)], [nb[j] for j in range(k)]
