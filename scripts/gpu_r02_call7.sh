timeout 240 python - <<'PY'
import sys, time
sys.path.insert(0, '.')
from rivulus_b200 import capi
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)]
ctx.profile_enable(True)
ctx.set_option(capi.OPT_CHUNK_PLAN, 1)
for n in (8_000_000, 64_000_000, 500_000_000):
    t = ctx.gen_batch(spec, n, 3_500_000_000)
    for thr in (499, 899, 998):
        for r in range(2):
            t0 = time.time()
            o = ctx.filter_project(t, capi.predicate(0, ">", thr), [0, 1, 2])
            ms = ctx.profile_read_launches()
            print("chunk plan rows", n, "thr", thr, "survivors", o.num_rows(), "device ms", [round(x, 3) for x in ms], "wall", round(time.time() - t0, 3), flush=True)
            o.release()
    t.release()
PY
echo "rc=$?"
