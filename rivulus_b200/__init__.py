"""rivulus_b200 — B200-native (sm_100a) filter / project / limit path of CleConor/rivulus.

Layout (only what the hot path needs):
  csrc/     hand-written CUDA kernels + the extern "C" ABI declared in include/rivulus_gpu.h
  host/     C++17 host layer mirroring the reference API (LazyFrame, Expr, DataFrame, RecordBatch, DataStream)
  capi.py   ctypes binding of the C ABI (tests, bench)
  sharding.py  row-range sharding helpers for the one-process-per-GPU launch

No CPU fallback exists: data-path calls need the built library and a B200.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
__version__ = "0.1.0"
