"""Host-side cost of one rvl_stream_push + rvl_stream_next for tiny batches (pure API / launch overhead, no PCIe time to speak of)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rivulus_b200 import capi

ctx = capi.Context(0)
rng = np.random.default_rng(1)
for rows in (1024, 65536):
    n_b = 512
    n = rows * n_b
    k, kb = capi.pinned_like(rng.integers(0, 1000, n).astype(np.int64))
    a, ab = capi.pinned_like(rng.integers(-2**62, 2**62, n).astype(np.int64))
    b, bb = capi.pinned_like(rng.random(n) * 1000.0)
    structs = []
    for i in range(n_b):
        cs = [capi.Column(capi.INT64, rows, i * rows, k), capi.Column(capi.INT64, rows, i * rows, a), capi.Column(capi.FLOAT64, rows, i * rows, b)]
        structs.append(((capi.RvlColumn * 3)(*[c.as_struct() for c in cs]), cs))
    for transfer in (capi.TRANSFER_STAGED, capi.TRANSFER_ZERO_COPY):
        best = 1e9
        for rep in range(4):
            st = ctx.open_stream([capi.INT64, capi.INT64, capi.FLOAT64], capi.predicate(0, ">", 899), [1, 2], -1, rows, 3, transfer)
            ctx.synchronize()
            t0 = time.perf_counter()
            inflight = 0
            for arr, _ in structs:
                st.push_structs(arr, 3)
                inflight += 1
                if inflight >= 2:
                    st.next_batch().release(); inflight -= 1
            while inflight:
                st.next_batch().release(); inflight -= 1
            ctx.synchronize()
            best = min(best, (time.perf_counter() - t0) / n_b * 1e6)
            st.close()
        print(f"rows/batch {rows:6d} transfer {transfer}: {best:7.1f} us per batch (push + next)")
