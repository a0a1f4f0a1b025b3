
RVL_CHUNK_DEBUG=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunk" > gpurun_out/r02_pytest_chunk9.log 2>&1; echo "pytest-chunk rc=$?"; tail -3 gpurun_out/r02_pytest_chunk9.log
cat > /tmp/c5s.py <<'PY'
import sys, os
sys.path.insert(0, '.')
from rivulus_b200 import capi
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)]
ctx.profile_enable(True)
t = ctx.gen_batch(spec, 500_000_000, 3_500_000_000)
for plan in (2, 0):
    ctx.set_option(capi.OPT_CHUNK_PLAN, plan)
    for thr in (998, 899, 799, 699, 499, 99):
        ms = []
        for r in range(5):
            o = ctx.filter_project(t, capi.predicate(0, ">", thr), [0, 1, 2])
            ms += ctx.profile_read_launches()
            cs = [o.checksum(j) for j in range(3)] if r == 0 else cs
            o.release()
        print("plan", plan, "thr", thr, "device ms", [round(x, 3) for x in ms[1:]], "checksum", cs[0] % 1000, cs[1] % 1000, cs[2] % 1000, flush=True)
PY
RVL_CHUNK_DEBUG=2 timeout 60 python /tmp/c5s.py; echo "timing rc=$?"
