"""One rvl_hash_join_inner at the bench.py `join` size (16 M x 64 M rows) — the command line ncu lists the launches of."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rivulus_b200 import capi
nb, npr, span = 16_000_000, 64_000_000, 20_000_000
rng = np.random.default_rng(11)
ctx = capi.Context(0)
build = ctx.upload([capi.Column(capi.INT64, nb, 0, rng.permutation(nb).astype(np.int64)), capi.Column(capi.INT64, nb, 0, np.arange(nb, dtype=np.int64))])
probe = ctx.upload([capi.Column(capi.INT64, npr, 0, rng.integers(0, span, npr).astype(np.int64)), capi.Column(capi.INT64, npr, 0, np.arange(npr, dtype=np.int64))])
for r in range(2):
    o = ctx.hash_join_inner(build, 0, probe, 0, [1], [1]); print(o.num_rows()); o.release()
