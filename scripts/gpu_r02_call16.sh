RVL_CHUNK_DEBUG=2 timeout 100 python scripts/chunk_hang_probe3.py; echo "probe rc=$?"
RVL_CHUNK_DEBUG=2 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunk_plan or two_pass" > gpurun_out/r02_pytest_chunk6.log 2>&1; echo "pytest-chunk rc=$?"; tail -6 gpurun_out/r02_pytest_chunk6.log
timeout 120 python - <<'PY'
import sys, time, os
sys.path.insert(0, '.')
from rivulus_b200 import capi
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)]
ctx.profile_enable(True)
t = ctx.gen_batch(spec, 500_000_000, 3_500_000_000)
for chunk in (0, 1):
    ctx.set_option(capi.OPT_CHUNK_PLAN, chunk)
    for thr in (998, 899, 799, 499, 99):
        for r in range(3):
            o = ctx.filter_project(t, capi.predicate(0, ">", thr), [0, 1, 2])
            ms = ctx.profile_read_launches()
            o.release()
        print("c5 shard chunk_plan", chunk, "thr", thr, "device ms", [round(x, 3) for x in ms], flush=True)
PY
