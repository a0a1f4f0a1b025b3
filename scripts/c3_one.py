"""One configs[2] RecordBatch (50 M rows {x: Float64, name: String, v: Int64}, 10 % nulls) through filter(x > T).select([name, v]) —
the command line ncu profiles for the string kernels.  --kernel 1|2 picks the round-1 / round-2 kernel pair."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rivulus_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=50_000_000)
ap.add_argument("--kernel", type=int, default=2)
ap.add_argument("--dense-min", type=int, default=None)
ap.add_argument("--lits", default="900,500")
ap.add_argument("--reps", type=int, default=1)
args = ap.parse_args()
ctx = capi.Context(0)
ctx.set_option(capi.OPT_STRING_KERNEL, args.kernel)
if args.dense_min is not None:
    ctx.set_option(capi.OPT_STRING_DENSE_MIN, args.dense_min)
t = ctx.gen_batch([(capi.SYNTH_F64, 0, 10), (capi.SYNTH_STR, 1, 10), (capi.SYNTH_I64, 2, 10)], args.rows, 0)
ctx.profile_enable(True)
for _ in range(args.reps):
    for lit in [float(x) for x in args.lits.split(",")]:
        out = ctx.filter_project(t, capi.predicate(0, ">", lit), [1, 2])
        print(lit, out.num_rows(), "device ms", [round(x, 3) for x in ctx.profile_read_launches()])
        out.release()
