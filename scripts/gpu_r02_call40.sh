timeout 600 python -m pytest tests/test_host_golden.py tests/test_gpu_parity.py -m gpu -x -q -k "join" 2>&1 | tail -4
python - <<'PY'
import time, sys
sys.path.insert(0, '.')
import numpy as np
from rivulus_b200 import capi
nb, npr, span = 16_000_000, 64_000_000, 20_000_000
rng = np.random.default_rng(11)
ctx = capi.Context(0)
bkeys = rng.permutation(nb).astype(np.int64); pkeys = rng.integers(0, span, npr).astype(np.int64)
build = ctx.upload([capi.Column(capi.INT64, nb, 0, bkeys), capi.Column(capi.INT64, nb, 0, np.arange(nb, dtype=np.int64))])
probe = ctx.upload([capi.Column(capi.INT64, npr, 0, pkeys), capi.Column(capi.INT64, npr, 0, np.arange(npr, dtype=np.int64))])
for r in range(4):
    ctx.synchronize(); t0 = time.perf_counter()
    o = ctx.hash_join_inner(build, 0, probe, 0, [1], [1]); ctx.synchronize()
    print("join 16M x 64M: %.2f ms, pairs %d" % ((time.perf_counter() - t0) * 1e3, o.num_rows()), flush=True); o.release()
# heavy keys: 1000 distinct keys on both sides (runs of 16 K build rows): the gallop path
bk2 = (bkeys % 1000).astype(np.int64); pk2 = (pkeys[:200_000] % 1000).astype(np.int64)
b2 = ctx.upload([capi.Column(capi.INT64, nb, 0, bk2)]); p2 = ctx.upload([capi.Column(capi.INT64, 200_000, 0, pk2)])
t0 = time.perf_counter(); o = ctx.hash_join_inner(b2, 0, p2, 0, [0], []); ctx.synchronize()
print("heavy keys: %.2f ms, pairs %d (expected %d)" % ((time.perf_counter() - t0) * 1e3, o.num_rows(), 200_000 * 16_000))
PY
