"""CPU-only checks of the drop-in boundary and the host logic: the C-ABI library loads and exports every symbol
include/rivulus_gpu.h declares (no compute calls), the host-side sharding arithmetic, the synthetic generator's
host/device-shared definition, and the roofline arithmetic bench.py reports."""
import ctypes
import os
import re

import numpy as np
import pytest

from rivulus_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rivulus_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rvl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    declared = header_symbols()
    assert len(declared) >= 35
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in include/rivulus_gpu.h but not exported: {missing}"
    assert sorted(capi.ABI_SYMBOLS) == declared, "capi.ABI_SYMBOLS and the header disagree"
    assert lib.rvl_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes the C compiler gives the header's structs (compiled here with gcc) must equal the ctypes mirrors
    import subprocess
    import tempfile
    src = '#include "rivulus_gpu.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(rvl_column), sizeof(rvl_predicate), sizeof(rvl_stream_config));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    assert sizes == [ctypes.sizeof(capi.RvlColumn), ctypes.sizeof(capi.RvlPredicate), ctypes.sizeof(capi.RvlStreamConfig)]


def test_no_gpu_is_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.RivulusError) as ei:
        capi.Context(0)
    assert ei.value.status == capi.CUDA


def test_product_never_touches_the_oracle():
    # the oracle is test infrastructure: nothing under rivulus_b200/ may import, include or link it
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rivulus_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"(?m)^\s*(from|import)\s+oracle\b|#include\s*[\"<][^\">]*oracle|liboracle", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_shard_range_partitions_rows():
    for n in (0, 1, 63, 64, 65, 1000, 10**9, 4 * 10**9, 4 * 10**9 + 17):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                b, e = capi.shard_range(n, r, world)
                assert b == prev and b <= e <= n
                if e < n:
                    assert e % 64 == 0          # bitmap words never straddle two GPUs
                prev = e
            assert prev == n
    assert capi.shard_range(4 * 10**9, 3, 8) == (1_500_000_000, 2_000_000_000)


def test_shard_limit_split():
    assert capi.shard_limit_split([5, 7, 9], -1) == [5, 7, 9]
    assert capi.shard_limit_split([5, 7, 9], 0) == [0, 0, 0]
    assert capi.shard_limit_split([5, 7, 9], 6) == [5, 1, 0]
    assert capi.shard_limit_split([5, 7, 9], 12) == [5, 7, 0]
    assert capi.shard_limit_split([5, 7, 9], 100) == [5, 7, 9]
    assert capi.shard_limit_split([0, 0, 3], 2) == [0, 0, 2]


def test_generator_spec_matches_numpy_restatement():
    # include/rivulus_synth.h restated in numpy: pins the generator so host, oracle and device agree
    from oracle import oracle as O
    G = np.uint64(0x9E3779B97F4A7C15)

    def splitmix(x):
        with np.errstate(over="ignore"):
            x = x + G
            z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    n, row0 = 5000, 123456789
    rows = np.arange(row0, row0 + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        u = [splitmix(np.uint64(42) + np.uint64(c) * G + rows) for c in range(3)]
    df = O.DataFrame.synth([("k", capi.SYNTH_KEY1000, 0, 0), ("a", capi.SYNTH_I64, 1, 0), ("b", capi.SYNTH_F64, 2, 0)], n, row0)
    assert df.column("k") == [int(x) for x in (u[0] % np.uint64(1000))]
    assert df.column("a") == [int(x) for x in u[1].view(np.int64)]
    assert df.column("b") == [float(x) for x in ((u[2] >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0) * 1000.0)]
    sel = float(np.mean((u[0] % np.uint64(1000)) > 499))
    assert abs(sel - 0.5) < 0.03


def test_roofline_arithmetic_matches_baseline_table():
    # BASELINE.md §3 / SURVEY.md §8(d): B_alg for N = 1e9, 1 predicate + 4 projected 8-byte columns
    import bench
    want = {0.001: 8.16e9, 0.10: 22.21e9, 0.50: 54.00e9, 0.90: 68.80e9}
    for s, b in want.items():
        assert abs(bench.b_alg(10**9, s) - b) / b < 2e-3, (s, bench.b_alg(10**9, s))
