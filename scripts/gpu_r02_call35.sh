cat > /tmp/c5s.py <<'PY'
import sys, os
sys.path.insert(0, '.')
from rivulus_b200 import capi
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)]
ctx.profile_enable(True)
t = ctx.gen_batch(spec, 500_000_000, 3_500_000_000)
ctx.set_option(capi.OPT_CHUNK_PLAN, 2)
for thr in (799, 499, 99):
    ms = []
    for r in range(6):
        o = ctx.filter_project(t, capi.predicate(0, ">", thr), [0, 1, 2])
        ms += ctx.profile_read_launches()
        cs = [o.checksum(j) for j in range(3)] if r == 0 else cs
        o.release()
    print("nap", os.environ.get("RVL_CHUNK_NAP", "0"), "thr", thr, "device ms", [round(x, 3) for x in ms[1:]], "checksums", cs[0] % 1000, flush=True)
PY
for nap in 0 20 50 100 200 400; do RVL_CHUNK_NAP=$nap timeout 120 python /tmp/c5s.py; done
