"""Shared helpers for the GPU parity tests: build the SAME columns for the CUDA path (through the C ABI)
and for the CPU oracle, run the same query on both, and compare every output buffer bit for bit."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from oracle import oracle as O
from rivulus_b200 import capi

DT = {"null": capi.NULL, "bool": capi.BOOLEAN, "i64": capi.INT64, "f64": capi.FLOAT64, "str": capi.STRING}


class Col:
    """One test column described by raw numpy buffers (whole buffers + an (offset, length) window)."""

    def __init__(self, dtype: str, n_total: int, values=None, valid: Optional[np.ndarray] = None, strings: Optional[List[bytes]] = None,
                 offset: int = 0, length: Optional[int] = None):
        self.dtype = dtype
        self.n_total = n_total
        self.offset = offset
        self.length = n_total - offset if length is None else length
        self.valid = None if valid is None else np.asarray(valid, dtype=bool)
        self.values = values
        self.strings = strings
        if dtype == "str":
            lens = np.array([len(s) for s in strings], dtype=np.int64)
            self.str_offsets = np.zeros(n_total + 1, dtype=np.int32)
            self.str_offsets[1:] = np.cumsum(lens)
            self.str_data = np.frombuffer(b"".join(strings), dtype=np.uint8).copy() if n_total else np.zeros(0, np.uint8)

    # ---- GPU side
    def gpu(self) -> capi.Column:
        vbits = None if self.valid is None else capi.pack_bits(self.valid)
        if self.dtype == "i64":
            return capi.Column(capi.INT64, self.length, self.offset, np.ascontiguousarray(self.values, dtype=np.int64), vbits)
        if self.dtype == "f64":
            return capi.Column(capi.FLOAT64, self.length, self.offset, np.ascontiguousarray(self.values, dtype=np.float64), vbits)
        if self.dtype == "bool":
            return capi.Column(capi.BOOLEAN, self.length, self.offset, capi.pack_bits(self.values), vbits)
        if self.dtype == "str":
            return capi.Column(capi.STRING, self.length, self.offset, None, vbits, self.str_offsets, self.str_data)
        return capi.Column(capi.NULL, self.length, self.offset)

    # ---- oracle side
    def oracle(self) -> O.Array:
        vbits = None if self.valid is None else O.pack_bits(self.valid)
        if self.dtype == "i64":
            return O.Array.i64(self.values, vbits, self.offset, self.length)
        if self.dtype == "f64":
            return O.Array.f64(self.values, vbits, self.offset, self.length)
        if self.dtype == "bool":
            return O.Array.boolean(O.pack_bits(self.values), self.n_total, vbits, self.offset, self.length)
        if self.dtype == "str":
            return O.Array.string(self.str_offsets, self.str_data, vbits, self.offset, self.length)
        return O.Array.null(self.length)


def random_col(rng, dtype: str, n: int, null_frac: float = 0.0, offset: int = 0, tail: int = 0, **kw) -> Col:
    """n visible rows; `offset` hidden rows in front and `tail` behind (exercises sliced views)."""
    total = offset + n + tail
    valid = None if null_frac <= 0 else rng.random(total) >= null_frac
    if dtype == "i64":
        lo, hi = kw.get("lo", -1000), kw.get("hi", 1000)
        return Col("i64", total, rng.integers(lo, hi, total).astype(np.int64), valid, offset=offset, length=n)
    if dtype == "f64":
        v = rng.random(total) * kw.get("scale", 1000.0)
        if kw.get("specials"):
            idx = rng.integers(0, max(total, 1), max(total // 16, 1))
            v[idx] = rng.choice([np.nan, -0.0, 0.0, np.inf, -np.inf, 499.5], len(idx))
        return Col("f64", total, v, valid, offset=offset, length=n)
    if dtype == "bool":
        return Col("bool", total, rng.random(total) < kw.get("p_true", 0.5), valid, offset=offset, length=n)
    if dtype == "str":
        maxlen = kw.get("maxlen", 12)
        alphabet = kw.get("alphabet", b"abc")
        lens = rng.integers(0, maxlen + 1, total)
        raw = rng.integers(0, len(alphabet), int(lens.sum()))
        chars = np.frombuffer(alphabet, dtype=np.uint8)[raw].tobytes()
        strings, pos = [], 0
        for L in lens:
            strings.append(chars[pos:pos + L]); pos += L
        return Col("str", total, None, valid, strings, offset=offset, length=n)
    return Col("null", total, offset=offset, length=n)


def upload(ctx: capi.Context, cols: Sequence[Col]) -> capi.Batch:
    return ctx.upload([c.gpu() for c in cols])


def oracle_batch(cols: Sequence[Col]) -> O.RecordBatch:
    return O.RecordBatch.try_new([f"c{i}" for i in range(len(cols))], [c.oracle() for c in cols])


def assert_batches_equal(got: capi.Batch, want: O.RecordBatch, ctx=""):
    """Bit-exact comparison of a device batch with an oracle batch: dtypes, row count, null counts, values buffers
    (compared as raw 64-bit patterns, so NaN payloads and -0.0 count), bitmap presence and bytes, string offsets/bytes."""
    assert got.num_columns() == want.num_columns(), f"{ctx}: column count {got.num_columns()} vs {want.num_columns()}"
    assert got.num_rows() == want.num_rows(), f"{ctx}: rows {got.num_rows()} vs {want.num_rows()}"
    n = want.num_rows()
    for j in range(want.num_columns()):
        g = got.download_column(j)
        w = want.column(j)
        where = f"{ctx} column {j}"
        assert g.dtype == w.dtype, f"{where}: dtype {g.dtype} vs {w.dtype}"
        assert g.length == w.length == n, f"{where}: length {g.length} vs {w.length}"
        assert w.offset == 0
        assert g.null_count == w.null_count, f"{where}: null_count {g.null_count} vs {w.null_count}"
        nb = (n + 7) // 8
        if w.dtype in (capi.INT64, capi.FLOAT64):
            assert np.array_equal(g.values.view(np.uint64), w.values[:n].view(np.uint64)), f"{where}: values differ"
        elif w.dtype == capi.BOOLEAN:
            assert np.array_equal(g.values[:nb], w.values[:nb]), f"{where}: boolean value bitmap differs"
        elif w.dtype == capi.STRING:
            assert np.array_equal(g.offsets, w.offsets), f"{where}: string offsets differ"
            assert np.array_equal(g.data, w.data), f"{where}: string bytes differ"
        assert (g.validity is None) == (w.validity is None), f"{where}: validity bitmap presence differs (got {g.validity is not None})"
        if w.validity is not None:
            assert np.array_equal(g.validity[:nb], w.validity[:nb]), f"{where}: validity bitmap differs"


def run_cmp(ctx: capi.Context, cols: Sequence[Col], pred_col: int, op: str, literal, proj: Sequence[int], limit: int = -1, tag=""):
    gb = upload(ctx, cols)
    got = ctx.filter_project(gb, capi.predicate(pred_col, op, literal), proj, limit)
    want = oracle_batch(cols).filter_project_cmp(pred_col, op, literal, proj, limit)
    assert_batches_equal(got, want, f"{tag} [{op} {literal!r} limit={limit}]")
    return got


def run_mask(ctx: capi.Context, cols: Sequence[Col], mask_col: int, proj: Sequence[int], limit: int = -1, tag=""):
    gb = upload(ctx, cols)
    got = ctx.filter_project(gb, capi.mask_predicate(mask_col), proj, limit)
    want = oracle_batch(cols).filter_project_mask(mask_col, proj, limit)
    assert_batches_equal(got, want, f"{tag} [mask limit={limit}]")
    return got
