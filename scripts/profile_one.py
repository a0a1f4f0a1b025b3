"""One small fused filter+project launch per selectivity — the command line ncu profiles (see profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rivulus_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=50_000_000)
ap.add_argument("--thresholds", default="998,899,499,99")
ap.add_argument("--quiet-opts", action="store_true")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--plan", default="auto", choices=["auto", "fused", "two_pass"])
ap.add_argument("--sparse-max", type=int, default=None)
ap.add_argument("--dense-slots", type=int, default=None)
ap.add_argument("--dense-ctas", type=int, default=None)
ap.add_argument("--scan-warps", type=int, default=None)
ap.add_argument("--dense-warps", type=int, default=None)
ap.add_argument("--scan-slots", type=int, default=None)
args = ap.parse_args()
ctx = capi.Context(0)
ctx.set_option(capi.OPT_PLAN, {"auto": capi.PLAN_AUTO, "fused": capi.PLAN_FUSED, "two_pass": capi.PLAN_TWO_PASS}[args.plan])
for opt, val in ((capi.OPT_SPARSE_MAX, args.sparse_max), (capi.OPT_DENSE_SLOTS, args.dense_slots), (capi.OPT_DENSE_CTAS_PER_SM, args.dense_ctas),
                     (capi.OPT_SCAN_WARPS, args.scan_warps), (capi.OPT_DENSE_WARPS, args.dense_warps), (capi.OPT_SCAN_SLOTS, args.scan_slots)):
    if val is not None:
        ctx.set_option(opt, val)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_I64, 1, 0), (capi.SYNTH_F64, 2, 0), (capi.SYNTH_I64, 3, 0), (capi.SYNTH_F64, 4, 0)]
t = ctx.gen_batch(spec, args.rows)
ctx.profile_enable(True)
for _ in range(args.reps):
    for thr in [int(x) for x in args.thresholds.split(",")]:
        out = ctx.filter_project(t, capi.predicate(0, ">", thr), [1, 2, 3, 4])
        ms = ctx.profile_read_launches()
        print(thr, out.num_rows(), "device ms", [round(x, 3) for x in ms])
        out.release()
