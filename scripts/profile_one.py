"""One small fused filter+project launch per selectivity — the command line ncu profiles (see profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rivulus_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=50_000_000)
ap.add_argument("--thresholds", default="998,899,499,99")
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_I64, 1, 0), (capi.SYNTH_F64, 2, 0), (capi.SYNTH_I64, 3, 0), (capi.SYNTH_F64, 4, 0)]
t = ctx.gen_batch(spec, args.rows)
for _ in range(args.reps):
    for thr in [int(x) for x in args.thresholds.split(",")]:
        out = ctx.filter_project(t, capi.predicate(0, ">", thr), [1, 2, 3, 4])
        print(thr, out.num_rows())
        out.release()
