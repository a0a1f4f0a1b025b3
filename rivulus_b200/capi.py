"""ctypes binding of include/rivulus_gpu.h (librivulus_gpu.so) — the same entry points a Rust `-sys`
crate would bind.  Thin by design: numpy arrays in, numpy arrays out, every call goes through the C ABI.

There is no CPU fallback: importing works without a GPU (so the ABI can be inspected), but every
data-path call needs the built library and a B200; a missing library raises ImportError loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "librivulus_gpu.so")

# rvl_status
OK, COLUMN_NOT_FOUND, TYPE_MISMATCH, INVALID_OPERATION, LENGTH_MISMATCH, SCHEMA_MISMATCH, OUT_OF_BOUNDS, OFFSET_OVERFLOW, \
    CUDA, INVALID_ARGUMENT, OUT_OF_MEMORY = range(11)
STATUS_NAMES = ["OK", "COLUMN_NOT_FOUND", "TYPE_MISMATCH", "INVALID_OPERATION", "LENGTH_MISMATCH", "SCHEMA_MISMATCH",
                "OUT_OF_BOUNDS", "OFFSET_OVERFLOW", "CUDA", "INVALID_ARGUMENT", "OUT_OF_MEMORY"]
# rvl_dtype (execution/schema.rs:1-8)
NULL, BOOLEAN, INT64, FLOAT64, STRING = range(5)
DTYPE_NAMES = ["Null", "Boolean", "Int64", "Float64", "String"]
# rvl_op (expressions/expr.rs:15-29)
OPS = {"+": 0, "-": 1, "*": 2, "/": 3, "==": 4, "!=": 5, "<": 6, ">": 7, "<=": 8, ">=": 9, "and": 10, "or": 11}
HOST, DEVICE = 0, 1
PRED_CMP_LITERAL, PRED_BOOL_COLUMN, PRED_TRUE = 0, 1, 2
# include/rivulus_synth.h
SYNTH_KEY1000, SYNTH_I64, SYNTH_F64, SYNTH_BOOL, SYNTH_AGE100, SYNTH_STR = range(6)


class RvlColumn(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("location", C.c_int32), ("length", C.c_int64), ("offset", C.c_int64),
                ("values", C.c_void_p), ("validity", C.c_void_p), ("offsets", C.c_void_p), ("data", C.c_void_p),
                ("data_len", C.c_int64), ("null_count", C.c_int64)]


class RvlPredicate(C.Structure):
    _fields_ = [("mode", C.c_int32), ("column", C.c_int32), ("op", C.c_int32), ("lit_dtype", C.c_int32),
                ("lit_i64", C.c_int64), ("lit_f64", C.c_double), ("lit_str", C.c_char_p), ("lit_str_len", C.c_int64),
                ("lit_bool", C.c_int32), ("tag_column", C.c_int32)]


class RvlStreamConfig(C.Structure):
    _fields_ = [("batch_rows", C.c_int64), ("n_staging", C.c_int32), ("transfer", C.c_int32)]


TRANSFER_AUTO, TRANSFER_STAGED, TRANSFER_ZERO_COPY = 0, 1, 2

IPC_HANDLE_BYTES = 64


class RvlGatherHandle(C.Structure):
    """rvl_gather_handle: the CUDA IPC handles of one destination column (plain bytes, shipped between ranks as they are)."""
    _fields_ = [("dtype", C.c_int32), ("has_validity", C.c_int32), ("rows", C.c_int64), ("data_bytes", C.c_int64),
                ("values", C.c_uint8 * IPC_HANDLE_BYTES), ("validity", C.c_uint8 * IPC_HANDLE_BYTES),
                ("offsets", C.c_uint8 * IPC_HANDLE_BYTES), ("data", C.c_uint8 * IPC_HANDLE_BYTES)]


PLAN_AUTO, PLAN_FUSED, PLAN_TWO_PASS = 0, 1, 2
OPT_PLAN, OPT_TWO_PASS_MIN_ROWS, OPT_SPARSE_MAX, OPT_DENSE_SLOTS, OPT_DENSE_CTAS_PER_SM, OPT_SCAN_SLOTS, OPT_SCAN_WARPS, OPT_DENSE_WARPS, \
    OPT_BITS_OVERLAP, OPT_EXACT_ALLOC, OPT_STRING_KERNEL, OPT_STRING_DENSE_MIN, OPT_CHUNK_PLAN, OPT_SCAN_ITEM_ROWS = range(14)


class RivulusError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"[{STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else status}] {message}")
        self.status = status
        self.message = message


# every symbol include/rivulus_gpu.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "rvl_abi_version", "rvl_last_error", "rvl_device_count",
    "rvl_ctx_create", "rvl_ctx_destroy", "rvl_ctx_synchronize", "rvl_ctx_trim", "rvl_ctx_pool_stats", "rvl_ctx_cuda_stream", "rvl_ctx_device", "rvl_ctx_launch_count",
    "rvl_ctx_profile_enable", "rvl_ctx_profile_read", "rvl_ctx_profile_read_launches", "rvl_ctx_set_option",
    "rvl_host_alloc", "rvl_host_free",
    "rvl_batch_upload", "rvl_batch_wrap_device", "rvl_batch_release", "rvl_batch_num_rows", "rvl_batch_num_columns",
    "rvl_batch_column", "rvl_batch_download_column", "rvl_batch_count_true", "rvl_boolean_op", "rvl_batch_slice", "rvl_batch_select", "rvl_batch_take", "rvl_batch_concat",
    "rvl_hash_join_inner",
    "rvl_filter_project", "rvl_predicate_mask", "rvl_filter_project_launch", "rvl_filter_project_finish",
    "rvl_stream_open", "rvl_stream_push", "rvl_stream_flush", "rvl_stream_next", "rvl_stream_limit_reached", "rvl_stream_collect",
    "rvl_stream_stats", "rvl_stream_launches", "rvl_stream_close",
    "rvl_shard_range", "rvl_shard_limit_split", "rvl_filter_project_sharded", "rvl_gather_to",
    "rvl_gather_dest_create", "rvl_gather_dest_open", "rvl_gather_push", "rvl_gather_dest_finish",
    "rvl_gen_batch", "rvl_batch_checksum",
]

_lib = None


def lib():
    """Load librivulus_gpu.so.  Raises ImportError if it has not been built — never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(or `make -C rivulus_b200/csrc`).  rivulus_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.rvl_last_error.restype = C.c_char_p
        _lib = L
    return _lib


def check(rc: int):
    if rc != OK:
        raise RivulusError(rc, lib().rvl_last_error().decode(errors="replace"))


def _addr(a: Optional[np.ndarray]):
    return C.c_void_p(a.ctypes.data) if a is not None and a.size > 0 else C.c_void_p(a.ctypes.data if a is not None else None)


# ------------------------------------------------------------------------------------------ host columns
@dataclass
class Column:
    """Host-side Arrow-layout column (numpy buffers): the argument/result type of the binding."""
    dtype: int
    length: int
    offset: int = 0
    values: Optional[np.ndarray] = None    # int64 / float64 values, or uint8 value-bitmap bytes (Boolean)
    validity: Optional[np.ndarray] = None  # uint8 bitmap bytes, or None (no nulls)
    offsets: Optional[np.ndarray] = None   # int32
    data: Optional[np.ndarray] = None      # uint8
    null_count: int = -1

    # ---- constructors from logical values (None = null), reference-builder layout
    @staticmethod
    def from_list(vals: Sequence, dtype: int) -> "Column":
        n = len(vals)
        valid = np.array([v is not None for v in vals], dtype=bool)
        validity = None if valid.all() or dtype == NULL else pack_bits(valid)
        if dtype == INT64:
            return Column(INT64, n, 0, np.array([0 if v is None else v for v in vals], dtype=np.int64), validity)
        if dtype == FLOAT64:
            return Column(FLOAT64, n, 0, np.array([0.0 if v is None else v for v in vals], dtype=np.float64), validity)
        if dtype == BOOLEAN:
            return Column(BOOLEAN, n, 0, pack_bits([bool(v) for v in vals]), validity)
        if dtype == STRING:
            enc = [b"" if v is None else (v.encode() if isinstance(v, str) else v) for v in vals]
            off = np.zeros(n + 1, np.int32)
            if n:
                off[1:] = np.cumsum([len(e) for e in enc])
            return Column(STRING, n, 0, None, validity, off, np.frombuffer(b"".join(enc), dtype=np.uint8).copy())
        return Column(NULL, n)

    def to_list(self) -> list:
        def bit(buf, i):
            return (int(buf[i >> 3]) >> (i & 7)) & 1
        out = []
        for r in range(self.length):
            li = self.offset + r
            if self.dtype == NULL or (self.validity is not None and not bit(self.validity, li)):
                out.append(None)
            elif self.dtype == INT64: out.append(int(self.values[li]))
            elif self.dtype == FLOAT64: out.append(float(self.values[li]))
            elif self.dtype == BOOLEAN: out.append(bool(bit(self.values, li)))
            else: out.append(self.data[self.offsets[li]:self.offsets[li + 1]].tobytes().decode())
        return out

    def as_struct(self) -> RvlColumn:
        s = RvlColumn()
        s.dtype, s.location, s.length, s.offset = self.dtype, HOST, self.length, self.offset
        s.values = self.values.ctypes.data if self.values is not None else None
        s.validity = self.validity.ctypes.data if self.validity is not None else None
        s.offsets = self.offsets.ctypes.data if self.offsets is not None else None
        s.data = self.data.ctypes.data if self.data is not None and self.data.size else None
        s.data_len = int(self.data.size) if self.data is not None else 0
        s.null_count = -1
        return s


def pack_bits(bools) -> np.ndarray:
    b = np.asarray(bools, dtype=np.uint8)
    return np.packbits(b, bitorder="little") if b.size else np.zeros(0, np.uint8)


def unpack_bits(buf: np.ndarray, n: int, offset: int = 0) -> np.ndarray:
    return np.unpackbits(buf, bitorder="little")[offset:offset + n].astype(bool)


def predicate(column: int, op: str, literal) -> RvlPredicate:
    """PhysicalPlan::Filter { column, value, op } (physical_plan/plan.rs:17-22)."""
    p = RvlPredicate()
    p.mode, p.column, p.op = PRED_CMP_LITERAL, column, OPS[op]
    if literal is None:
        p.lit_dtype = NULL
    elif isinstance(literal, (bool, np.bool_)):
        p.lit_dtype, p.lit_bool = BOOLEAN, int(bool(literal))
    elif isinstance(literal, (int, np.integer)):
        p.lit_dtype, p.lit_i64 = INT64, int(literal)
    elif isinstance(literal, (float, np.floating)):
        p.lit_dtype, p.lit_f64 = FLOAT64, float(literal)
    elif isinstance(literal, (str, bytes)):
        b = literal.encode() if isinstance(literal, str) else literal
        p.lit_dtype, p.lit_str, p.lit_str_len = STRING, b, len(b)
        p._keep = b
    else:
        raise TypeError(type(literal))
    return p


def mask_predicate(column: int) -> RvlPredicate:
    """FilterStream { predicate_column } (execution/stream.rs:117-120)."""
    p = RvlPredicate()
    p.mode, p.column = PRED_BOOL_COLUMN, column
    return p


def true_predicate() -> RvlPredicate:
    p = RvlPredicate()
    p.mode = PRED_TRUE
    return p


# ------------------------------------------------------------------------------------------ handles
class Context:
    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().rvl_ctx_create(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().rvl_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(lib().rvl_ctx_synchronize(self._h))

    def pool_stats(self):
        """(reserved, used) bytes of the device memory pool."""
        r, u = C.c_uint64(0), C.c_uint64(0)
        check(lib().rvl_ctx_pool_stats(self._h, C.byref(r), C.byref(u)))
        return r.value, u.value

    def trim(self):
        """rvl_ctx_trim: give the pool's cached blocks back to the driver (between workloads of different footprints)."""
        check(lib().rvl_ctx_trim(self._h))

    def cuda_stream(self) -> int:
        s = C.c_void_p()
        check(lib().rvl_ctx_cuda_stream(self._h, C.byref(s)))
        return s.value or 0

    def launch_count(self) -> int:
        n = C.c_int64()
        check(lib().rvl_ctx_launch_count(self._h, C.byref(n)))
        return n.value

    def set_option(self, option: int, value: int):
        """rvl_ctx_set_option: OPT_PLAN (PLAN_AUTO / PLAN_FUSED / PLAN_TWO_PASS) and the two-pass tuning knobs."""
        check(lib().rvl_ctx_set_option(self._h, C.c_int32(option), C.c_int64(value)))

    def profile_enable(self, on: bool = True):
        check(lib().rvl_ctx_profile_enable(self._h, int(on)))

    def profile_read(self):
        """(summed device ms of the fused kernel, launches) since the last read."""
        ms, n = C.c_double(), C.c_int64()
        check(lib().rvl_ctx_profile_read(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_read_launches(self, cap: int = 1 << 16) -> List[float]:
        """Per-launch device times (ms) of the fused kernel, in launch order, since the last read."""
        buf = (C.c_double * cap)()
        n = C.c_int64()
        check(lib().rvl_ctx_profile_read_launches(self._h, buf, C.c_int64(cap), C.byref(n)))
        return [buf[i] for i in range(min(n.value, cap))]

    # ---- batches
    def upload(self, cols: Sequence[Column]) -> "Batch":
        arr = (RvlColumn * max(len(cols), 1))(*[c.as_struct() for c in cols])
        out = C.c_void_p()
        check(lib().rvl_batch_upload(self._h, arr, len(cols), C.byref(out)))
        return Batch(self, out)

    def wrap_device(self, views: Sequence[RvlColumn]) -> "Batch":
        """rvl_batch_wrap_device: a batch over caller-owned DEVICE buffers (e.g. column views of other batches), no copy.
        The caller keeps the owning batches alive."""
        arr = (RvlColumn * max(len(views), 1))(*views)
        out = C.c_void_p()
        check(lib().rvl_batch_wrap_device(self._h, arr, len(views), C.byref(out)))
        b = Batch(self, out)
        b._keep = list(views)
        return b

    def gen_batch(self, cols: Sequence[tuple], n: int, row0: int = 0) -> "Batch":
        """cols: [(kind, col_id, null_pct)] — synthetic columns generated on the device (rivulus_synth.h)."""
        kinds = (C.c_int32 * len(cols))(*[c[0] for c in cols])
        ids = (C.c_uint32 * len(cols))(*[c[1] for c in cols])
        nulls = (C.c_uint32 * len(cols))(*[c[2] for c in cols])
        out = C.c_void_p()
        check(lib().rvl_gen_batch(self._h, kinds, ids, nulls, len(cols), C.c_uint64(row0), C.c_int64(n), C.byref(out)))
        return Batch(self, out)

    def concat(self, batches: Sequence["Batch"]) -> "Batch":
        arr = (C.c_void_p * max(len(batches), 1))(*[b._h.value for b in batches])
        out = C.c_void_p()
        check(lib().rvl_batch_concat(self._h, arr, len(batches), C.byref(out)))
        return Batch(self, out)

    # ---- the hot path
    def filter_project(self, batch: "Batch", pred: Optional[RvlPredicate], proj: Sequence[int], limit: int = -1) -> "Batch":
        p = (C.c_int32 * max(len(proj), 1))(*proj)
        out = C.c_void_p()
        check(lib().rvl_filter_project(self._h, batch._h, C.byref(pred) if pred is not None else None, p, len(proj),
                                       C.c_int64(limit), C.byref(out)))
        return Batch(self, out)

    def filter_project_launch(self, batch: "Batch", pred, proj: Sequence[int], limit: int = -1):
        p = (C.c_int32 * max(len(proj), 1))(*proj)
        out = C.c_void_p()
        check(lib().rvl_filter_project_launch(self._h, batch._h, C.byref(pred) if pred is not None else None, p, len(proj),
                                              C.c_int64(limit), C.byref(out)))
        return out

    def filter_project_finish(self, pending) -> "Batch":
        out = C.c_void_p()
        check(lib().rvl_filter_project_finish(self._h, pending, C.byref(out)))
        return Batch(self, out)

    def hash_join_inner(self, build: "Batch", build_key: int, probe: "Batch", probe_key: int, probe_proj: Sequence[int], build_proj: Sequence[int],
                        build_tag_column: int = 0, probe_tag_column: int = 0) -> "Batch":
        """rvl_hash_join_inner (physical_plan/plan.rs:174-284): probe_proj columns then build_proj columns, one row per matching pair."""
        pp = (C.c_int32 * max(len(probe_proj), 1))(*probe_proj)
        bp = (C.c_int32 * max(len(build_proj), 1))(*build_proj)
        out = C.c_void_p()
        n = C.c_int64(0)
        check(lib().rvl_hash_join_inner(self._h, build._h, build_key, build_tag_column, probe._h, probe_key, probe_tag_column,
                                        pp, len(probe_proj), bp, len(build_proj), C.byref(out), C.byref(n)))
        return Batch(self, out)

    def predicate_mask(self, batch: "Batch", pred: RvlPredicate) -> "Batch":
        out = C.c_void_p()
        check(lib().rvl_predicate_mask(self._h, batch._h, C.byref(pred), C.byref(out)))
        return Batch(self, out)

    def boolean_op(self, op: str, a: "Batch", a_col: int, b: Optional["Batch"] = None, b_col: int = 0) -> "Batch":
        """BooleanArray::and / or / not (array/boolean.rs:120-165) on device columns; returns a one-column batch."""
        out = C.c_void_p()
        check(lib().rvl_boolean_op(self._h, {"and": 0, "or": 1, "not": 2}[op], a._h, a_col, b._h if b is not None else None, b_col, C.byref(out)))
        return Batch(self, out)

    def open_stream(self, dtypes: Sequence[int], pred, proj: Sequence[int], limit: int = -1, batch_rows: int = 1 << 20,
                    n_staging: int = 2, transfer: int = TRANSFER_AUTO) -> "Stream":
        return Stream(self, dtypes, pred, proj, limit, batch_rows, n_staging, transfer)


class Batch:
    """Device-resident RecordBatch handle (execution/record_batch.rs:8-13)."""

    def __init__(self, ctx: Context, handle):
        self.ctx = ctx
        self._h = handle

    def release(self):
        if self._h:
            lib().rvl_batch_release(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def num_rows(self) -> int:
        n = C.c_int64()
        check(lib().rvl_batch_num_rows(self._h, C.byref(n)))
        return n.value

    def num_columns(self) -> int:
        n = C.c_int32()
        check(lib().rvl_batch_num_columns(self._h, C.byref(n)))
        return n.value

    def view(self, i: int) -> RvlColumn:
        v = RvlColumn()
        check(lib().rvl_batch_column(self._h, i, C.byref(v)))
        return v

    def count_true(self, i: int) -> int:
        """BooleanArray::count_true of column i."""
        n = C.c_int64()
        check(lib().rvl_batch_count_true(self.ctx._h, self._h, i, C.byref(n)))
        return n.value

    def slice(self, offset: int, length: int) -> "Batch":
        out = C.c_void_p()
        check(lib().rvl_batch_slice(self._h, C.c_int64(offset), C.c_int64(length), C.byref(out)))
        return Batch(self.ctx, out)

    def take(self, indices: Sequence[int]) -> "Batch":
        """RecordBatch::take (record_batch.rs:108-129): rows by index."""
        a = np.ascontiguousarray(indices, dtype=np.int64)
        out = C.c_void_p()
        check(lib().rvl_batch_take(self.ctx._h, self._h, a.ctypes.data_as(C.POINTER(C.c_int64)) if a.size else None, C.c_int64(a.size), C.byref(out)))
        return Batch(self.ctx, out)

    def select(self, indices: Sequence[int]) -> "Batch":
        a = (C.c_int32 * max(len(indices), 1))(*indices)
        out = C.c_void_p()
        check(lib().rvl_batch_select(self._h, a, len(indices), C.byref(out)))
        return Batch(self.ctx, out)

    def checksum(self, i: int) -> int:
        c = C.c_uint64()
        check(lib().rvl_batch_checksum(self.ctx._h, self._h, i, C.byref(c)))
        return c.value

    def download_column(self, i: int) -> Column:
        """Copy column i to the host, rebased to offset 0 (what a freshly built reference array holds)."""
        v = self.view(i)
        n = v.length
        col = Column(v.dtype, n, 0, null_count=v.null_count)
        if v.dtype in (INT64, FLOAT64):
            col.values = np.zeros(n, dtype=np.int64 if v.dtype == INT64 else np.float64)
        elif v.dtype == BOOLEAN:
            col.values = np.zeros((n + 7) // 8, dtype=np.uint8)
        elif v.dtype == STRING:
            col.offsets = np.zeros(n + 1, dtype=np.int32)
            col.data = np.zeros(max(v.data_len, 1), dtype=np.uint8)
        if v.validity and v.dtype != NULL:
            col.validity = np.zeros((n + 7) // 8, dtype=np.uint8)
        s = col.as_struct()
        if col.data is not None:
            s.data = col.data.ctypes.data
            s.data_len = int(col.data.size)
        check(lib().rvl_batch_download_column(self.ctx._h, self._h, i, C.byref(s)))
        if col.data is not None:
            col.data = col.data[:s.data_len]
        col.null_count = s.null_count
        return col

    def download(self) -> List[Column]:
        return [self.download_column(i) for i in range(self.num_columns())]


class Stream:
    """trait DataStream over host batches (execution/stream.rs:25-54) with pinned, overlapped H2D."""

    def __init__(self, ctx: Context, dtypes, pred, proj, limit, batch_rows, n_staging, transfer=TRANSFER_AUTO):
        self.ctx = ctx
        self._h = C.c_void_p()
        d = (C.c_int32 * max(len(dtypes), 1))(*dtypes)
        p = (C.c_int32 * max(len(proj), 1))(*proj)
        cfg = RvlStreamConfig(batch_rows, n_staging, transfer)
        check(lib().rvl_stream_open(ctx._h, d, len(dtypes), C.byref(pred) if pred is not None else None, p, len(proj),
                                    C.c_int64(limit), C.byref(cfg), C.byref(self._h)))
        # page-locked sources are read asynchronously (copy engine, or in place by the kernels): the pushed columns are kept
        # alive here until the stream has handed out everything in flight (include/rivulus_gpu.h: rvl_stream_push)
        self._held = []
        self._pred = pred

    def push(self, cols: Sequence[Column]) -> bool:
        arr = (RvlColumn * max(len(cols), 1))(*[c.as_struct() for c in cols])
        acc = C.c_int32()
        check(lib().rvl_stream_push(self._h, arr, len(cols), C.byref(acc)))
        if acc.value:
            self._held.append((arr, list(cols)))
        return bool(acc.value)

    def push_structs(self, arr, n) -> bool:
        """Push a prebuilt rvl_column array; the caller keeps the buffers it points into alive."""
        acc = C.c_int32()
        check(lib().rvl_stream_push(self._h, arr, n, C.byref(acc)))
        return bool(acc.value)

    def flush(self):
        check(lib().rvl_stream_flush(self._h))

    def launches(self) -> int:
        n = C.c_int64()
        check(lib().rvl_stream_launches(self._h, C.byref(n)))
        return n.value

    def next_batch(self) -> Optional[Batch]:
        out = C.c_void_p()
        has = C.c_int32()
        check(lib().rvl_stream_next(self._h, C.byref(out), C.byref(has)))
        if not has.value:
            self._held.clear()   # nothing is in flight any more
        return Batch(self.ctx, out) if has.value else None

    def limit_reached(self) -> bool:
        r = C.c_int32()
        check(lib().rvl_stream_limit_reached(self._h, C.byref(r)))
        return bool(r.value)

    def collect(self) -> Batch:
        out = C.c_void_p()
        check(lib().rvl_stream_collect(self._h, C.byref(out)))
        self._held.clear()
        return Batch(self.ctx, out)

    def stats(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().rvl_stream_stats(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"batches_pushed": a.value, "batches_skipped": b.value, "h2d_bytes": c.value}

    def close(self):
        if self._h:
            lib().rvl_stream_close(self._h)
            self._h = C.c_void_p()
        self._held = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PinnedBuffer:
    """Page-locked host memory (rvl_host_alloc) exposed as a numpy array — the staging memory of the streaming path."""

    def __init__(self, nbytes: int):
        self._p = C.c_void_p()
        self.nbytes = max(int(nbytes), 1)
        check(lib().rvl_host_alloc(C.c_size_t(self.nbytes), C.byref(self._p)))
        self.u8 = np.frombuffer((C.c_uint8 * self.nbytes).from_address(self._p.value), dtype=np.uint8)

    def view(self, dtype, count=None):
        a = self.u8.view(dtype)
        return a if count is None else a[:count]

    def free(self):
        if self._p:
            self.u8 = None
            lib().rvl_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def pinned_like(a: np.ndarray):
    """Copy `a` into pinned memory; returns (array view, owner)."""
    buf = PinnedBuffer(a.nbytes)
    v = buf.view(a.dtype, a.size)
    v[...] = a.reshape(-1)
    return v, buf


def device_count() -> int:
    n = C.c_int32()
    rc = lib().rvl_device_count(C.byref(n))
    return n.value if rc == OK else 0


def shard_range(n_rows: int, rank: int, world: int):
    b, e = C.c_int64(), C.c_int64()
    check(lib().rvl_shard_range(C.c_int64(n_rows), rank, world, C.byref(b), C.byref(e)))
    return b.value, e.value


def shard_limit_split(counts: Sequence[int], limit: int):
    n = len(counts)
    c = (C.c_int64 * n)(*counts)
    t = (C.c_int64 * n)()
    check(lib().rvl_shard_limit_split(c, n, C.c_int64(limit), t))
    return list(t)


def filter_project_sharded(ctxs: Sequence[Context], shards: Sequence[Batch], pred, proj: Sequence[int], limit: int = -1):
    n = len(ctxs)
    ca = (C.c_void_p * n)(*[c._h.value for c in ctxs])
    sa = (C.c_void_p * n)(*[s._h.value for s in shards])
    p = (C.c_int32 * max(len(proj), 1))(*proj)
    outs = (C.c_void_p * n)()
    counts = (C.c_int64 * n)()
    check(lib().rvl_filter_project_sharded(ca, n, sa, C.byref(pred) if pred is not None else None, p, len(proj), C.c_int64(limit),
                                           outs, counts))
    return [Batch(ctxs[g], C.c_void_p(outs[g])) for g in range(n)], list(counts)


def gather_dest_create(ctx: Context, dtypes: Sequence[int], has_validity: Sequence[bool], total_rows: int, data_bytes: Optional[Sequence[int]] = None):
    """rvl_gather_dest_create: (destination batch, handle bytes to ship to the other ranks)."""
    n = len(dtypes)
    d = (C.c_int32 * max(n, 1))(*dtypes)
    v = (C.c_int32 * max(n, 1))(*[int(bool(x)) for x in has_validity])
    nb = (C.c_int64 * max(n, 1))(*([int(x) for x in data_bytes] if data_bytes is not None else [0] * n))
    handles = (RvlGatherHandle * max(n, 1))()
    out = C.c_void_p()
    check(lib().rvl_gather_dest_create(ctx._h, d, v, nb, n, C.c_int64(total_rows), C.byref(out), handles))
    return Batch(ctx, out), bytes(handles)[:C.sizeof(RvlGatherHandle) * n]


def gather_dest_open(ctx: Context, handle_bytes: bytes) -> Batch:
    """rvl_gather_dest_open: map another rank's destination buffers (peer access over NVLink)."""
    n = len(handle_bytes) // C.sizeof(RvlGatherHandle)
    handles = (RvlGatherHandle * max(n, 1)).from_buffer_copy(handle_bytes.ljust(C.sizeof(RvlGatherHandle) * max(n, 1), b"\0"))
    out = C.c_void_p()
    check(lib().rvl_gather_dest_open(ctx._h, handles, n, C.byref(out)))
    return Batch(ctx, out)


def gather_push(ctx: Context, part: Batch, dest: Batch, row_offset: int, byte_offsets: Optional[Sequence[int]] = None):
    """rvl_gather_push: write `part` into rows [row_offset, row_offset + part.num_rows()) of the destination."""
    bo = None
    if byte_offsets is not None:
        bo = (C.c_int64 * max(len(byte_offsets), 1))(*[int(x) for x in byte_offsets])
    check(lib().rvl_gather_push(ctx._h, part._h, dest._h, C.c_int64(row_offset), bo))


def gather_dest_finish(ctx: Context, dest: Batch):
    check(lib().rvl_gather_dest_finish(ctx._h, dest._h))


def gather_to(dst: Context, parts: Sequence[Batch]) -> Batch:
    """rvl_gather_to: ordered physical concatenation of per-GPU results on dst's device (peer reads over NVLink)."""
    arr = (C.c_void_p * max(len(parts), 1))(*[b._h.value for b in parts])
    out = C.c_void_p()
    check(lib().rvl_gather_to(dst._h, arr, len(parts), C.byref(out)))
    return Batch(dst, out)
