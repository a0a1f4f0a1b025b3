"""Measurements of the BASELINE.json configs that are not the bench.py headline:

  c3  configs[2]: 200 M rows {x: Float64, name: String (8..40 B, mean 24), v: Int64}, 10 % nulls in every column, in RecordBatches of
      <= 50 M rows (i32 string offsets), filter(x > T).select([name, v]) at 10 % / 50 %, plus one `<` run ("nulls pass")
  c5  configs[4], one shard: {k: Int64, v: Float64, f: Boolean}, filter(k > T).select([k, v, f]) at 10 % / 50 %
  c4  configs[3]: collect_streaming() shape — host-resident batches {k, a, b: 8 B, flag: Boolean} of 64 K / 256 K / 1 M rows, 64 batches,
      LIMIT 1000: (i) filter(flag) (reference-expressible) (ii) filter(k > T) at 0.1 % / 10 %; and the same stream without LIMIT
      (H2D rate against a plain pinned cudaMemcpy of the same bytes)

Every figure is device time from CUDA events on the library stream (rvl_ctx_profile_*), or host wall clock around a synchronised
region for the streaming runs.  Prints one JSON object.  Not a bench.py line: these are the per-config numbers DESIGN.md §5 quotes.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rivulus_b200 import capi  # noqa: E402

PEAK = 6535.7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def p_sector(s, w):
    return 1.0 - (1.0 - s) ** (32.0 / w)


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


def run_c3(ctx, rows, batch_rows, reps):
    """Strings + nulls.  B_alg per SURVEY.md §8(d): predicate column + its validity, per projected column the touched 32-B sectors
    (+ validity), survivor bytes read + written, new offsets/values/validity written."""
    out = []
    nb = (rows + batch_rows - 1) // batch_rows
    spec = [(capi.SYNTH_F64, 0, 10), (capi.SYNTH_STR, 1, 10), (capi.SYNTH_I64, 2, 10)]
    batches = [ctx.gen_batch(spec, min(batch_rows, rows - i * batch_rows), i * batch_rows) for i in range(nb)]
    in_str_bytes = sum(b.view(1).data_len for b in batches)
    for op, lit, label in ((">", 900.0, "10%"), (">", 500.0, "50%"), ("<", 100.0, "lt: 10% + nulls pass")):
        times, surv, sbytes = [], 0, 0
        for r in range(reps + 1):
            ctx.profile_read_launches()
            surv = sbytes = 0
            for b in batches:
                o = ctx.filter_project(b, capi.predicate(0, op, lit), [1, 2])
                surv += o.num_rows(); sbytes += o.view(0).data_len
                o.release()
            t = sum(ctx.profile_read_launches())
            if r > 0 or reps == 0:
                times.append(t)
        ms = median(times)
        s = surv / rows
        lbar = in_str_bytes / (rows * 0.9)          # mean length of a non-null string
        b_alg = rows * (8 + 1 / 8)                                                        # x + validity
        b_alg += rows * (4 * p_sector(s, 4) + (1 / 8) * p_sector(s, 1 / 8))              # name offsets + validity sectors
        b_alg += sbytes + surv * 4 + sbytes + surv / 8                                    # survivor bytes read; offsets, bytes, validity written
        b_alg += rows * (8 * p_sector(s, 8) + (1 / 8) * p_sector(s, 1 / 8)) + surv * (8 + 1 / 8)   # v
        out.append({"query": f"filter(x {op} {lit}).select([name, v])", "label": label, "rows": rows, "batches": nb, "survivors": surv,
                    "selectivity": s, "survivor_string_bytes": sbytes, "mean_len": lbar, "device_ms": ms, "rows_per_s": rows / ms * 1e3,
                    "b_alg_gb": b_alg / 1e9, "alg_gbs": b_alg / ms / 1e6, "frac_of_peak": b_alg / ms / 1e6 / PEAK})
    for b in batches:
        b.release()
    return out


def run_c5(ctx, rows, reps):
    out = []
    spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)]
    t = ctx.gen_batch(spec, rows, 3_000_000_000)
    for thr, label in ((899, "10%"), (499, "50%")):
        times, surv = [], 0
        for r in range(reps + 1):
            ctx.profile_read_launches()
            o = ctx.filter_project(t, capi.predicate(0, ">", thr), [0, 1, 2])
            surv = o.num_rows(); o.release()
            tt = sum(ctx.profile_read_launches())
            if r > 0 or reps == 0:
                times.append(tt)
        ms = median(times)
        s = surv / rows
        b_alg = rows * 8 + rows * 8 * p_sector(s, 8) + rows / 8 * p_sector(s, 1 / 8) + surv * (16 + 1 / 8)
        out.append({"query": f"filter(k > {thr}).select([k, v, f])", "label": label, "rows": rows, "survivors": surv, "device_ms": ms,
                    "rows_per_s": rows / ms * 1e3, "b_alg_gb": b_alg / 1e9, "alg_gbs": b_alg / ms / 1e6, "frac_of_peak": b_alg / ms / 1e6 / PEAK})
    t.release()
    return out


def run_c4(ctx, n_batches, reps, transfer=0):
    out = []
    rng = np.random.default_rng(7)
    for batch_rows in (65536, 262144, 1048576):
        n = batch_rows * n_batches
        k, kb = capi.pinned_like(rng.integers(0, 1000, n).astype(np.int64))
        a, ab = capi.pinned_like(rng.integers(-2**62, 2**62, n).astype(np.int64))
        b, bb = capi.pinned_like(rng.random(n) * 1000.0)
        f, fb = capi.pinned_like(np.packbits(rng.integers(0, 2, n).astype(np.uint8), bitorder="little"))
        dtypes = [capi.INT64, capi.INT64, capi.FLOAT64, capi.BOOLEAN]
        batch_bytes = batch_rows * 24 + batch_rows // 8

        def cols(i):
            o = i * batch_rows
            return [capi.Column(capi.INT64, batch_rows, o, k), capi.Column(capi.INT64, batch_rows, o, a),
                    capi.Column(capi.FLOAT64, batch_rows, o, b), capi.Column(capi.BOOLEAN, batch_rows, o, f)]

        out_pin = [capi.PinnedBuffer(n * 8), capi.PinnedBuffer(n * 8)]
        structs = []
        for i in range(n_batches):
            cs = cols(i)
            structs.append(((capi.RvlColumn * 4)(*[c.as_struct() for c in cs]), cs))
        queries = [("filter(flag).select([k,a]).limit(1000)", capi.mask_predicate(3), [0, 1], 1000),
                   ("filter(k > 998).select([a,b]).limit(1000)", capi.predicate(0, ">", 998), [1, 2], 1000),
                   ("filter(k > 899).select([a,b]).limit(1000)", capi.predicate(0, ">", 899), [1, 2], 1000),
                   ("filter(k > 899).select([a,b])  (no limit)", capi.predicate(0, ">", 899), [1, 2], -1)]
        for label, pred, proj, limit in queries:
            walls, stats, rows_out = [], None, 0
            for r in range(reps + 1):
                st = ctx.open_stream(dtypes, pred, proj, limit, batch_rows, 3, transfer)
                ctx.synchronize()
                t0 = time.perf_counter()
                for arr, _keep in structs:
                    if not st.push_structs(arr, 4):
                        break
                res = st.collect()
                rows_out = res.num_rows()
                nres = rows_out
                for j in range(len(proj)):                                 # D2H of the result (pinned destination) inside the timed region
                    sct = capi.Column(capi.INT64 if res.view(j).dtype == capi.INT64 else capi.FLOAT64, nres, 0,
                                      out_pin[j].view(np.int64 if res.view(j).dtype == capi.INT64 else np.float64, max(nres, 1))).as_struct()
                    capi.check(capi.lib().rvl_batch_download_column(ctx._h, res._h, j, C.byref(sct)))
                got = None
                ctx.synchronize()
                t1 = time.perf_counter()
                stats = st.stats()
                st.close(); res.release(); del got
                if r > 0 or reps == 0:
                    walls.append((t1 - t0) * 1e3)
            ms = median(walls)
            ideal = None
            if limit >= 0:
                # batches a perfect LimitStream pulls: until the running survivor count reaches the limit
                kk = np.asarray(k)
                if pred.mode == capi.PRED_BOOL_COLUMN:
                    keep = np.unpackbits(np.asarray(f), bitorder="little")[:n].astype(bool)
                else:
                    keep = kk > pred.lit_i64
                cum = np.cumsum(keep)
                idx = int(np.searchsorted(cum, limit))
                ideal = min(n_batches, idx // batch_rows + 1)
            out.append({"batch_rows": batch_rows, "n_batches": n_batches, "query": label, "rows_out": rows_out, "wall_ms": ms,
                        "batches_transferred": stats["batches_pushed"], "batches_ideal": ideal, "h2d_bytes": stats["h2d_bytes"],
                        "h2d_gbs": stats["h2d_bytes"] / ms / 1e6, "stream_bytes_total": batch_bytes * n_batches})
        del structs
        for x in (kb, ab, bb, fb, *out_pin):
            x.free()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c3,c5,c4")
    ap.add_argument("--c3-rows", type=int, default=200_000_000)
    ap.add_argument("--c3-batch-rows", type=int, default=50_000_000)
    ap.add_argument("--c5-rows", type=int, default=500_000_000, help="one shard of the 4 B-row table at 8 GPUs")
    ap.add_argument("--c4-batches", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--plan", default="auto", choices=["auto", "fused", "two_pass"])
    ap.add_argument("--transfer", default="auto", choices=["auto", "staged", "zero_copy"])
    args = ap.parse_args()
    ctx = capi.Context(0)
    ctx.set_option(capi.OPT_PLAN, {"auto": capi.PLAN_AUTO, "fused": capi.PLAN_FUSED, "two_pass": capi.PLAN_TWO_PASS}[args.plan])
    ctx.profile_enable(True)
    res = {"peak_gbs": PEAK, "plan": args.plan}
    which = args.which.split(",")
    if "c3" in which:
        res["c3"] = run_c3(ctx, args.c3_rows, args.c3_batch_rows, args.reps)
    if "c5" in which:
        res["c5"] = run_c5(ctx, args.c5_rows, args.reps)
    if "c4" in which:
        ctx.profile_enable(False)
        res["c4"] = run_c4(ctx, args.c4_batches, max(2, args.reps // 2), {"auto": 0, "staged": 1, "zero_copy": 2}[args.transfer])
        res["c4_transfer"] = args.transfer
    print(json.dumps(res))


if __name__ == "__main__":
    main()
