for dbg in 0 1; do
RVL_CHUNK_DEBUG=$dbg timeout 120 python - <<'PY'
import sys, time, os
sys.path.insert(0, '.')
from rivulus_b200 import capi
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)]
ctx.profile_enable(True)
t = ctx.gen_batch(spec, 500_000_000, 3_500_000_000)
for thr in (499, 998):
    for r in range(3):
        o = ctx.filter_project(t, capi.predicate(0, ">", thr), [0, 1])
        ms = ctx.profile_read_launches()
        o.release()
    print("debug", os.environ.get("RVL_CHUNK_DEBUG"), "thr", thr, "device ms", [round(x, 3) for x in ms], flush=True)
PY
done
