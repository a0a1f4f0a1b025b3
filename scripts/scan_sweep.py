"""predicate_scan_kernel: warps per CTA x ring slots per warp, 1e9-row Int64 column, 0.1 % query (scan + a tiny gather)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rivulus_b200 import capi
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_I64, 1, 0)]
t = ctx.gen_batch(spec, 1_000_000_000)
ctx.profile_enable(True)
import itertools
combos = [(8, 0, d) for d in (1, 2, 3)] + [(8, 512, d) for d in (2, 3, 4, 5, 6)] + [(8, 256, d) for d in (4, 6, 8, 10, 12)] + \
         [(16, 0, d) for d in (1, 2, 3)] + [(16, 256, d) for d in (2, 3, 4, 6)] + [(32, 0, d) for d in (1, 2)]
if len(sys.argv) > 1:
    combos = [(8, 0, 2), (8, 0, 3), (8, 512, 3), (16, 0, 2), (16, 256, 2)]
for warps, rows, slots in combos:
    if True:
        ctx.set_option(capi.OPT_SCAN_WARPS, warps); ctx.set_option(capi.OPT_SCAN_SLOTS, slots); ctx.set_option(capi.OPT_SCAN_ITEM_ROWS, rows)
        ms = []
        for r in range(8):
            o = ctx.filter_project(t, capi.predicate(0, ">", 1000), [1]); o.num_rows(); o.release()   # no survivor: the scan alone
            ms += ctx.profile_read_launches()
        ms = sorted(ms[2:])
        print(f"W={warps:2d} R={rows or 8192 // warps:4d} D={slots:2d} ({warps * slots * (rows or 8192 // warps) * 8 // 1024:3d} KB/SM)  scan-only invocation: min {ms[0]:.4f} ms  median {ms[len(ms)//2]:.4f} ms  -> {8.0 / ms[len(ms)//2]:.2f} TB/s", flush=True)
