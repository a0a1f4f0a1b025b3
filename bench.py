#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native filter / project / limit path.

Workload (BASELINE.json configs[1]): per GPU a 1 B-row table {k: Int64, a: Int64, b: Float64, c: Int64, d: Float64},
query `filter(k > T).select([a, b, c, d])` swept over T = 998 / 899 / 499 / 99 (selectivity 0.1 / 10 / 50 / 90 %).
One STEP = the four queries of the sweep, each one fused pass over the table.  `value` = input rows scanned per
second over the whole job with the table resident in HBM; `e2e` = the same metric through the streaming C ABI with
the table in pinned HOST memory (H2D of every batch and D2H of every result inside the timed region).

Multi-GPU (torchrun, one rank per GPU): weak scaling — rank g owns rows [g*R, (g+1)*R) of a G*R-row table
(row-range sharding); no data-path collective, torch.distributed only for the barrier and the
max-over-ranks of the device time.

Sub-records of the same JSON line (each bounded to a few seconds, each checked against the oracle / golden checksums):
  c5  BASELINE configs[4] at every N: the 4 B-row {k: Int64, v: Float64, f: Boolean} table row-range sharded over the N ranks
      (strong scaling), filter(k > T).select([k, v, f]) at 10 % / 50 %, per-rank device time, B_alg fraction, and the
      order-preserving physical concatenation onto rank 0 timed separately (rvl_gather_to at N = 1, the CUDA-IPC
      rvl_gather_* push over NVLink at N > 1); count + order-sensitive checksums vs tests/golden/synth_checksums.json
  c1  configs[0] (N = 1): 1 M-row {name, age}, LazyFrame.from_dataframe(df).filter(age > 25).select([name]).collect() through
      rivulus_b200.frame (host layer -> C ABI), wall clock, against the oracle's eager engine at the same 1 M rows on 1 core
  c3  configs[2] (N = 1): 200 M rows, String projection + 10 % nulls
  c4  configs[3] (N = 1): collect_streaming() shape, 64 K - 1 M-row host batches, LIMIT 1000 early stop, pinned H2D overlap

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--impl native|reference]
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

THRESHOLDS = [(998, 0.001), (899, 0.10), (499, 0.50), (99, 0.90)]
# the workload both arms run (BASELINE.json configs[1]); the same string goes into config.workload of both JSON lines
WORKLOAD = ("configs[1]: filter(k > T).select([a,b,c,d]) over {k,a:Int64,b:Float64,c:Int64,d:Float64}, "
            "T in 998/899/499/99 (0.1/10/50/90 %), 4 queries per step")
N_PROJ = 4
BYTES_PER_ROW_IN = 40  # 5 x 8-byte columns


def b_alg(n_rows: int, s: float, n_proj: int = N_PROJ, w: int = 8) -> float:
    """Algorithmic HBM bytes of one fused pass: predicate column + the 32-byte sectors of each
    projected column holding >= 1 survivor + the compacted output."""
    p_sector = 1.0 - (1.0 - s) ** (32 // w)
    return w * n_rows + n_proj * w * n_rows * p_sector + n_proj * s * n_rows * w


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml).  The thread is started (and its first,
    slow NVML queries made) before the region opens; only samples taken after arm() are kept."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, uuid=None, index=0, period=0.02):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self.armed = False          # samples count only once the timed region has opened (arm())
        self._stop_evt = threading.Event()
        self.period = period
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            self.h = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                watts = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                if self.armed:
                    self.samples.append(mhz)
                    for bit, name in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(name)
                    self.power.append(watts)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def arm(self):
        self.armed = True

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.ok or not self.samples:
            return None
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s), "power_w_max": max(self.power) if self.power else None}


# ------------------------------------------------------------------------------------------ CPU arms (oracle)
def cpu_eager_sweep(rows: int, threads: int, steps: int, warmup: int):
    """The reference algorithm (oracle/: C++ restatement of the eager engine) on `threads` row-range shards of
    `rows` rows each, timed over `steps` sweeps.  Returns (rows/s, total seconds)."""
    from oracle import oracle as O
    from rivulus_b200 import capi
    spec = [("k", capi.SYNTH_KEY1000, 0, 0), ("a", capi.SYNTH_I64, 1, 0), ("b", capi.SYNTH_F64, 2, 0),
            ("c", capi.SYNTH_I64, 3, 0), ("d", capi.SYNTH_F64, 4, 0)]
    dfs = [O.DataFrame.synth(spec, rows, row0=t * rows) for t in range(threads)]
    total = 0.0
    for it in range(warmup + steps):
        for thr, _ in THRESHOLDS:
            if threads == 1:
                secs, _ = O.time_eager_filter_select(dfs[0], "k", ">", thr, ["a", "b", "c", "d"])
            else:
                secs, _ = O.time_eager_filter_select_mt(dfs, "k", ">", thr, ["a", "b", "c", "d"])
            if it >= warmup:
                total += secs
    return rows * threads * len(THRESHOLDS) * steps / total, total


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = min(os.cpu_count() or 1, 64)
    rows = args.cpu_rows
    t0 = time.time()
    value, secs = cpu_eager_sweep(rows, threads, args.steps, min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": "rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": secs * 1000.0 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_gpu": args.rows, "global_rows": args.rows * max(args.gpus, 1),
                   "engine": "reference eager engine: LazyFrame.filter(..).select(..).collect() per query",
                   "note": "the reference is single-threaded Rust and cannot be compiled in this image (no rustc); this arm runs "
                           "oracle/ — the C++ restatement of its eager engine — as one instance per host core on row-range shards"},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port",
                         "sample": f"{threads} shards x {rows} rows x 4 queries x {args.steps} steps ({secs:.1f} s of wall time)"},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from rivulus_b200 import capi
    from rivulus_b200.sharding import shard_rows

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the native arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # optional (RVL_BENCH_AFFINITY=1): keep each rank (and the pinned host table it allocates for the e2e leg) on the NUMA node of
    # its own GPU, so that on a multi-socket host eight ranks do not pull their PCIe traffic through one socket
    numa = "unbound"
    if world > 1 and os.environ.get("RVL_BENCH_AFFINITY", "0") == "1":   # opt-in: no benefit measured on this pool's single-NUMA-node hosts
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = None
            for cand in (uuid, "GPU-" + uuid):
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode())
                    break
                except Exception:
                    h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            pynvml.nvmlDeviceSetCpuAffinity(h)
            numa = f"bound to the GPU's CPU set ({len(os.sched_getaffinity(0))} cores)"
        except Exception as e:   # best effort: affinity is an optimisation, never a requirement
            numa = f"unbound ({type(e).__name__})"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    ctx = capi.Context(local_rank)
    ctx.set_option(capi.OPT_PLAN, {"auto": capi.PLAN_AUTO, "fused": capi.PLAN_FUSED, "two_pass": capi.PLAN_TWO_PASS}[args.plan])
    for opt, val in ((capi.OPT_SPARSE_MAX, args.sparse_max), (capi.OPT_DENSE_SLOTS, args.dense_slots), (capi.OPT_DENSE_CTAS_PER_SM, args.dense_ctas),
                     (capi.OPT_SCAN_WARPS, args.scan_warps), (capi.OPT_DENSE_WARPS, args.dense_warps), (capi.OPT_SCAN_SLOTS, args.scan_slots)):
        if val is not None:
            ctx.set_option(opt, val)
    rows = args.rows
    begin, end = shard_rows(rows * world, rank, world)
    assert end - begin == rows, (begin, end, rows)
    table_spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_I64, 1, 0), (capi.SYNTH_F64, 2, 0), (capi.SYNTH_I64, 3, 0), (capi.SYNTH_F64, 4, 0)]
    table = ctx.gen_batch(table_spec, rows, begin)
    preds = [capi.predicate(0, ">", thr) for thr, _ in THRESHOLDS]
    proj = [1, 2, 3, 4]

    trace = []

    def step(tag=None):
        counts = []
        for p in preds:
            if tag is not None and args.trace:
                t0 = time.perf_counter(); r0 = ctx.pool_stats()
            out = ctx.filter_project(table, p, proj)   # kernels + count readback (AUTO plan: scan + compaction passes at this size)
            counts.append(out.num_rows())
            if tag is not None and args.trace:
                trace.append((tag, round((time.perf_counter() - t0) * 1e3, 3), r0[0] >> 20, ctx.pool_stats()[0] >> 20))
            out.release()
        return counts

    # ---- parity spot-check inside the bench (oracle = checker only): first rows of this shard
    parity = "skipped"
    if args.verify_rows > 0 and rank == 0:
        from oracle import oracle as O
        vr = min(args.verify_rows, rows)
        sl = table.slice(0, vr)
        for thr, _ in THRESHOLDS[:3:2]:
            out = ctx.filter_project(sl, capi.predicate(0, ">", thr), proj)
            cnt, sums = O.synth_filter_checksums(vr, begin, capi.SYNTH_KEY1000, 0, ">", thr, [(s[0], s[1]) for s in table_spec[1:]])
            if out.num_rows() != cnt or [out.checksum(j) for j in range(4)] != sums:
                raise SystemExit(f"bench.py: GPU result differs from the oracle at T={thr}")
            out.release()
        parity = f"count+checksums == oracle on first {vr} rows"

    for _ in range(max(args.warmup, 3)):
        counts = step()
    warm = max(args.warmup, 3)

    stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=torch.device("cuda", local_rank))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(uuid=str(torch.cuda.get_device_properties(local_rank).uuid), index=local_rank)
    # the sampler's first NVML queries take the driver lock for milliseconds: let them happen before the timed region opens
    # (one 13 ms outlier in the very first timed operator was exactly that), then keep sampling through it
    sampler.start()
    time.sleep(0.25)
    ctx.profile_enable(True)
    step(-1)                    # one more untimed step with profiling on (event creation paths warm)
    ctx.profile_read_launches()
    barrier(); torch.cuda.synchronize()
    launches0 = ctx.launch_count()
    sampler.arm()
    e0.record(stream)
    for i_step in range(args.steps):
        counts = step(i_step)
    e1.record(stream)
    e1.synchronize()
    clocks = sampler.stop()
    if args.trace and rank == 0:
        print("trace (step, host ms, pool reserved MiB before -> after):", trace[:12], file=sys.stderr)
    torch.cuda.synchronize(); barrier()
    ms_total = e0.elapsed_time(e1)
    gpu_launches = ctx.launch_count() - launches0
    per_launch = ctx.profile_read_launches()
    ctx.profile_enable(False)

    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_launches = torch.tensor([gpu_launches], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(total_launches, op=dist.ReduceOp.SUM)

    # ---- roofline of the operator's kernels, from the per-invocation events of the timed region
    peak, peak_src = load_peaks()
    sweep, alg_total, kern_total = [], 0.0, 0.0
    for qi, (thr, s_nom) in enumerate(THRESHOLDS):
        times = per_launch[qi::len(THRESHOLDS)]
        s_act = counts[qi] / rows
        alg = b_alg(rows, s_act)
        avg_ms = sum(times) / max(len(times), 1)
        alg_total += alg * len(times)
        kern_total += sum(times)
        st = sorted(times) or [0.0]
        sweep.append({"threshold": thr, "selectivity": s_act, "survivors": counts[qi], "kernel_ms": avg_ms,
                      "kernel_ms_min_median_max": [st[0], st[len(st) // 2], st[-1]],
                      "kernel_ms_per_step": [round(x, 4) for x in times], "b_alg_gb": alg / 1e9,
                      "alg_gbs": alg / 1e9 / (avg_ms / 1e3) if avg_ms > 0 else None,
                      "frac_of_peak": alg / 1e9 / (avg_ms / 1e3) / peak if avg_ms > 0 else None,
                      "rows_per_s": rows / (avg_ms / 1e3) if avg_ms > 0 else None,
                      "b_scan_gb": (BYTES_PER_ROW_IN * rows + N_PROJ * 8 * counts[qi]) / 1e9})
    achieved = alg_total / 1e9 / (kern_total / 1e3) if kern_total > 0 else 0.0
    # dram__bytes_read + dram__bytes_write per operator invocation from the committed `ncu --set full` capture (profiles/traffic.json):
    # a STATIC figure measured once on one GPU at this workload, not re-measured by this run — labelled so, and left out at N > 1
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and rows == 1_000_000_000 and os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("dram_bytes_per_launch")
            traffic_src = "static, not re-measured by this run: " + str(tj.get("source", "profiles/traffic.json"))
        except Exception:
            traffic = None

    value = rows * world * len(THRESHOLDS) * args.steps / (ms_max / 1e3)
    line = {
        "metric": "rows_per_sec", "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "engine": "one rvl_filter_project call per query", "plan": args.plan,
                   "rows_per_gpu": rows, "global_rows": rows * world, "partitioning": f"row-range x{world}", "host_affinity": numa,
                   "l2": "inputs (40 B/row x rows) far exceed the 126 MB L2; no flush needed",
                   "timing": "CUDA events on the library stream around K steps incl. count readback; max over ranks"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "kernel": ("fused_filter_project_kernel<kPredI64>" if args.plan == "fused" else
                                "predicate_scan_kernel<kPredI64> + compact_dense_kernel + gather_sparse_kernel (one operator invocation)"),
                     "definition": "sum of algorithmic bytes (SURVEY 8(d): 8.16/22.2/54.0/68.8 B per row at 0.1/10/50/90 %, x rows) of the timed "
                                   "operator invocations / sum of their device durations (CUDA events on the library stream around the kernels "
                                   "of each invocation)"},
        "sweep": sweep, "gpu_launches": int(total_launches.item()), "clocks": clocks, "parity": parity,
    }

    # ---- full-size parity (rank 0 owns rows [0, R)): survivor count + order-sensitive checksum of every projected column of
    # every query against the values the CPU oracle computed from the generator (tests/golden/synth_checksums.json)
    if rank == 0 and not args.no_golden:
        line["parity_full"] = golden_check(ctx, table, rows, begin, preds, proj)

    # ---- CPU baseline (rank 0, every N): the oracle's eager engine on a bounded sample, 1 core like the reference
    if rank == 0 and args.cpu_rows > 0:
        v, secs = cpu_eager_sweep(args.cpu_rows, 1, 1, 0)
        line["cpu_baseline"] = {"value": v, "unit": "rows/s", "cores": 1, "kind": "port",
                                "sample": f"{args.cpu_rows} rows x 4 queries, eager collect() restatement ({secs:.1f} s)"}
    barrier()

    # ---- end to end through the streaming C ABI with HOST buffers
    if not args.no_e2e:
        line["e2e"] = run_e2e(args, ctx, table, preds, proj, rank, world, local_rank, barrier)

    # ---- the other BASELINE configs as sub-records (each guarded: a failure is reported in the record, not by losing the line)
    table.release()
    del table
    ctx.trim()
    if not args.no_configs:
        peak, _ = load_peaks()
        line["c5"] = guarded(run_c5, args, ctx, rank, world, local_rank, barrier, peak)
        if world == 1:
            ctx.trim()
            line["c3"] = guarded(run_c3, args, ctx, peak)
            ctx.trim()
            line["c4"] = guarded(run_c4, args, ctx)
            line["c1"] = guarded(run_c1, args)
            line["csv"] = guarded(run_csv, args)
            line["join"] = guarded(run_join, args, ctx)

    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def guarded(fn, *a):
    try:
        return fn(*a)
    except Exception as e:   # noqa: BLE001 — the record says what failed; the headline line survives
        import traceback
        traceback.print_exc(file=sys.stderr)
        return {"error": f"{type(e).__name__}: {e}"}


def load_golden():
    try:
        with open(os.path.join(ROOT, "tests", "golden", "synth_checksums.json")) as f:
            return json.load(f)["cases"]
    except Exception:
        return []


def golden_case(workload, rows, row0, thr):
    for c in load_golden():
        if c["workload"] == workload and c["rows"] == rows and c["row0"] == row0 and c["pred"]["literal"] == thr:
            return c["count"], [int(x) for x in c["checksums"]]
    return None


def golden_check(ctx, table, rows, row0, preds, proj):
    from rivulus_b200 import capi  # noqa: F401
    checked = 0
    for (thr, _), p in zip(THRESHOLDS, preds):
        g = golden_case("configs[1]", rows, row0, thr)
        if g is None:
            continue
        out = ctx.filter_project(table, p, proj)
        got = (out.num_rows(), [out.checksum(j) for j in range(len(proj))])
        out.release()
        if got != (g[0], g[1]):
            raise SystemExit(f"bench.py: GPU result at {rows} rows, T={thr} differs from the oracle's golden count/checksums: {got} vs {g}")
        checked += 1
    if checked == 0:
        return f"no golden values for rows={rows} row0={row0}"
    return f"count + 4 column checksums == oracle golden at all {rows} rows for {checked}/4 queries"


def p_sector(s, w):
    return 1.0 - (1.0 - s) ** (32.0 / w)


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


# ------------------------------------------------------------------------------------------ c5: BASELINE configs[4]
C5_ROWS = 4_000_000_000
C5_THRESHOLDS = [(899, "10%"), (499, "50%")]


def run_c5(args, ctx, rank, world, local_rank, barrier, peak):
    """4 B-row {k, v, f: Boolean}, contiguous row ranges over the ranks (strong scaling), filter(k > T).select([k, v, f]),
    survivors concatenated in row order on rank 0 (record_batch.rs:245-342 across GPUs)."""
    import torch
    import torch.distributed as dist
    from rivulus_b200 import capi, sharding

    total_rows = args.c5_rows
    begin, end = sharding.shard_rows(total_rows, rank, world)
    n = end - begin
    spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)]
    table = ctx.gen_batch(spec, n, begin)
    dev = torch.device("cuda", local_rank)
    rec = {"workload": "configs[4]: filter(k > T).select([k, v, f]) over {k: Int64, v: Float64, f: Boolean}, row-range sharded, ordered concat on rank 0",
           "rows_total": total_rows, "rows_per_gpu": n, "n_gpus": world, "scaling": "strong", "queries": []}
    ctx.profile_enable(True)
    for thr, label in C5_THRESHOLDS:
        pred = capi.predicate(0, ">", thr)
        times, out = [], None
        for r in range(args.c5_reps + 1):
            if out is not None:
                out.release()
            ctx.profile_read_launches()
            barrier()
            out = ctx.filter_project(table, pred, [0, 1, 2])
            t = sum(ctx.profile_read_launches())
            if r > 0:
                times.append(t)
        ms = median(times)
        t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
        cnt = torch.tensor([out.num_rows()], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        ms_max, survivors = float(t_ms.item()), int(cnt.item())
        s_act = survivors / total_rows
        # k read once (predicate AND projected), v and f by 32-byte sector, survivors written
        b_alg = total_rows * 8 + total_rows * 8 * p_sector(s_act, 8) + total_rows / 8 * p_sector(s_act, 1 / 8) + survivors * (16 + 1 / 8)
        q = {"threshold": thr, "label": label, "survivors": survivors, "selectivity": s_act, "device_ms": ms_max,
             "rows_per_s": total_rows / (ms_max / 1e3), "b_alg_gb": b_alg / 1e9,
             "alg_gbs_per_gpu": b_alg / world / 1e9 / (ms_max / 1e3), "frac": b_alg / world / 1e9 / (ms_max / 1e3) / peak}

        # ---- ordered physical concatenation on rank 0, timed on its own
        marks = {}

        def timer(label_):
            ctx.synchronize()
            marks[label_] = time.perf_counter()
        if world == 1:
            # once untimed: the destination comes out of the stream-ordered pool, whose first growth to this size maps fresh memory
            capi.gather_to(ctx, [out]).release()
            ctx.synchronize()
            t0 = time.perf_counter()
            gathered = capi.gather_to(ctx, [out])
            ctx.synchronize()
            g_ms = (time.perf_counter() - t0) * 1e3
            how = "rvl_gather_to (one part: device-local concat, second call)"
        else:
            gathered = sharding.gather_ordered(ctx, out, 0, dev, timer)
            g_ms = (marks["push_end"] - marks["push_begin"]) * 1e3
            how = ("rvl_gather_dest_create/open/push/finish: CUDA IPC, every rank writes its rows into rank 0's memory over NVLink "
                   "(push phase only: the destination's cudaMalloc + handle exchange are outside the timed region)")
        t_g = torch.tensor([g_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_g, op=dist.ReduceOp.MAX)
        g_ms = float(t_g.item())
        g_bytes = survivors * 16 + survivors / 8
        q["gather"] = {"ms": g_ms, "bytes": g_bytes, "gbs": g_bytes / 1e9 / (g_ms / 1e3) if g_ms > 0 else None, "how": how,
                       "timing": "host wall clock between stream synchronisations around the push phase, max over ranks"}
        if rank == 0:
            gold = golden_case("configs[4]", total_rows, 0, thr)
            got = (gathered.num_rows(), [gathered.checksum(j) for j in range(3)])
            if gold is None:
                q["parity"] = f"no golden values for {total_rows} rows (count {got[0]})"
            elif got != (gold[0], gold[1]):
                raise SystemExit(f"bench.py: c5 gathered result at T={thr} differs from the oracle's golden count/checksums: {got} vs {gold}")
            else:
                q["parity"] = "gathered count + checksums of k, v, f (order-sensitive) == oracle golden at 4e9 rows"
            gathered.release()
        out.release()
        barrier()
        ctx.trim()
        rec["queries"].append(q)
    ctx.profile_enable(False)
    table.release()
    return rec


# ------------------------------------------------------------------------------------------ c3: BASELINE configs[2]
def run_c3(args, ctx, peak):
    """200 M rows {x: Float64, name: String (8..40 B), v: Int64}, 10 % nulls in every column, batches of <= 50 M rows (int32 string
    offsets), filter(x > T).select([name, v]) at 10 % / 50 % and one `<` run (nulls pass).  B_alg per SURVEY.md 8(d)."""
    from oracle import oracle as O
    from rivulus_b200 import capi
    rows, batch_rows, reps = args.c3_rows, 50_000_000, args.c3_reps
    nb = (rows + batch_rows - 1) // batch_rows
    spec = [(capi.SYNTH_F64, 0, 10), (capi.SYNTH_STR, 1, 10), (capi.SYNTH_I64, 2, 10)]
    batches = [ctx.gen_batch(spec, min(batch_rows, rows - i * batch_rows), i * batch_rows) for i in range(nb)]
    in_str_bytes = sum(b.view(1).data_len for b in batches)
    ctx.profile_enable(True)
    out = {"workload": "configs[2]: filter(x <op> T).select([name, v]) over {x: Float64, name: String, v: Int64}, 10 % nulls each",
           "rows": rows, "batches": nb, "queries": []}
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
            if r > 0:
                times.append(t)
        ms = median(times)
        s = surv / rows
        b_alg = rows * (8 + 1 / 8)                                                        # x + validity
        b_alg += rows * (4 * p_sector(s, 4) + (1 / 8) * p_sector(s, 1 / 8))              # name offsets + validity sectors
        b_alg += sbytes + surv * 4 + sbytes + surv / 8                                    # survivor bytes read; offsets, bytes, validity written
        b_alg += rows * (8 * p_sector(s, 8) + (1 / 8) * p_sector(s, 1 / 8)) + surv * (8 + 1 / 8)   # v
        out["queries"].append({"query": f"filter(x {op} {lit}).select([name, v])", "label": label, "survivors": surv, "selectivity": s,
                               "survivor_string_bytes": sbytes, "device_ms": ms, "rows_per_s": rows / ms * 1e3, "b_alg_gb": b_alg / 1e9,
                               "alg_gbs": b_alg / ms / 1e6, "frac": b_alg / ms / 1e6 / peak})
    # parity on a bounded prefix (the string / null checksum oracle is ~10 M rows/s): count, checksums, null counts, string bytes
    vr = min(args.c3_verify_rows, batches[0].num_rows())
    sl = batches[0].slice(0, vr)
    o = ctx.filter_project(sl, capi.predicate(0, ">", 500.0), [1, 2])
    cnt, sums, nulls, nbytes = O.synth_filter_checksums_nulls(vr, 0, (capi.SYNTH_F64, 0, 10), ">", 500.0, [(capi.SYNTH_STR, 1, 10), (capi.SYNTH_I64, 2, 10)])
    got = (o.num_rows(), [o.checksum(0), o.checksum(1)], [o.view(0).null_count, o.view(1).null_count], o.view(0).data_len)
    if got != (cnt, sums, nulls, nbytes[0]):
        raise SystemExit(f"bench.py: c3 GPU result differs from the oracle on the first {vr} rows: {got} vs {(cnt, sums, nulls, nbytes[0])}")
    out["parity"] = f"count, string + int checksums, null counts, string bytes == oracle on the first {vr} rows (50 %)"
    out["input_string_bytes"] = in_str_bytes
    o.release(); sl.release()
    ctx.profile_enable(False)
    for b in batches:
        b.release()
    return out


# ------------------------------------------------------------------------------------------ c4: BASELINE configs[3]
def pinned_copy_gbs(ctx, nbytes=256 << 20):
    """The box's pinned H2D rate, measured here and now (denominator of the c4 H2D figures)."""
    import numpy as np
    import torch
    from rivulus_b200 import capi
    buf = capi.PinnedBuffer(nbytes)
    arr = buf.view(np.int64)
    col = capi.Column(capi.INT64, arr.size, 0, arr)
    best = 0.0
    for _ in range(4):
        t0 = time.perf_counter()
        b = ctx.upload([col])
        dt = time.perf_counter() - t0
        b.release()
        best = max(best, nbytes / dt / 1e9)
    buf.free()
    return best


def run_c4(args, ctx):
    """collect_streaming() shape: a host-resident stream of {k, a, b: 8 B, flag: Boolean} batches of 64 K / 256 K / 1 M rows.
    LIMIT 1000 runs report the batches that crossed PCIe against what a LimitStream would pull (streaming.rs:269-271); the
    un-limited run reports the H2D rate (STAGED: every needed byte on the copy engine) and how much kernel time hid under it."""
    import ctypes as C
    import numpy as np
    from rivulus_b200 import capi
    rng = np.random.default_rng(7)
    n_batches, reps = args.c4_batches, 3
    slot_rows = 1 << 20
    copy_gbs = pinned_copy_gbs(ctx)
    out = {"workload": "configs[3]: host batches {k, a: Int64, b: Float64, flag: Boolean} -> rvl_stream_push/collect, LIMIT 1000 and un-limited",
           "pinned_copy_gbs": copy_gbs, "slot_rows": slot_rows, "n_batches": n_batches, "runs": []}
    for batch_rows in (65536, 262144, 1048576):
        n = batch_rows * n_batches
        k, kb = capi.pinned_like(rng.integers(0, 1000, n).astype(np.int64))
        a, ab = capi.pinned_like(rng.integers(-2**62, 2**62, n).astype(np.int64))
        b, bb = capi.pinned_like(rng.random(n) * 1000.0)
        f, fb = capi.pinned_like(np.packbits(rng.integers(0, 2, n).astype(np.uint8), bitorder="little"))
        dtypes = [capi.INT64, capi.INT64, capi.FLOAT64, capi.BOOLEAN]
        out_pin = [capi.PinnedBuffer(n * 8), capi.PinnedBuffer(n * 8)]
        structs = []
        for i in range(n_batches):
            o = i * batch_rows
            cs = [capi.Column(capi.INT64, batch_rows, o, k), capi.Column(capi.INT64, batch_rows, o, a),
                  capi.Column(capi.FLOAT64, batch_rows, o, b), capi.Column(capi.BOOLEAN, batch_rows, o, f)]
            structs.append(((capi.RvlColumn * 4)(*[c.as_struct() for c in cs]), cs))
        queries = [("filter(flag).select([k,a]).limit(1000)", capi.mask_predicate(3), [0, 1], 1000, capi.TRANSFER_AUTO),
                   ("filter(k > 998).select([a,b]).limit(1000)", capi.predicate(0, ">", 998), [1, 2], 1000, capi.TRANSFER_AUTO),
                   ("filter(k > 899).select([a,b]).limit(1000)", capi.predicate(0, ">", 899), [1, 2], 1000, capi.TRANSFER_AUTO),
                   ("filter(k > 899).select([a,b])  (no limit, STAGED)", capi.predicate(0, ">", 899), [1, 2], -1, capi.TRANSFER_STAGED),
                   ("filter(k > 899).select([a,b])  (no limit, AUTO)", capi.predicate(0, ">", 899), [1, 2], -1, capi.TRANSFER_AUTO)]
        for label, pred, proj, limit, transfer in queries:
            walls, streams, stats, rows_out, launches, kern_ms = [], [], None, 0, 0, 0.0
            for r in range(reps + 1):
                st = ctx.open_stream(dtypes, pred, proj, limit, slot_rows, 3, transfer)
                ctx.synchronize()
                ctx.profile_enable(True)
                t0 = time.perf_counter()
                for arr, _keep in structs:
                    if not st.push_structs(arr, 4):
                        break
                st.flush()
                ctx.synchronize()                                          # every H2D copy and every kernel of the stream has finished
                t_stream = time.perf_counter()
                res = st.collect()
                rows_out = res.num_rows()
                for j in range(len(proj)):                                 # D2H of the result (pinned destination) inside the timed region
                    dt = np.int64 if res.view(j).dtype == capi.INT64 else np.float64
                    sct = capi.Column(capi.INT64 if dt == np.int64 else capi.FLOAT64, rows_out, 0, out_pin[j].view(dt, max(rows_out, 1))).as_struct()
                    capi.check(capi.lib().rvl_batch_download_column(ctx._h, res._h, j, C.byref(sct)))
                ctx.synchronize()
                t1 = time.perf_counter()
                kern_ms = sum(ctx.profile_read_launches())
                ctx.profile_enable(False)
                stats, launches = st.stats(), st.launches()
                st.close(); res.release()
                if r > 0:
                    walls.append((t1 - t0) * 1e3)
                    streams.append((t_stream - t0) * 1e3)
            ms, stream_ms = median(walls), median(streams)
            rec = {"batch_rows": batch_rows, "query": label, "rows_out": rows_out, "wall_ms": ms, "batches_transferred": stats["batches_pushed"],
                   "operator_launches": launches, "h2d_copy_engine_bytes": stats["h2d_bytes"]}
            kk = np.asarray(k)
            keep = np.unpackbits(np.asarray(f), bitorder="little")[:n].astype(bool) if pred.mode == capi.PRED_BOOL_COLUMN else kk > pred.lit_i64
            if limit >= 0:
                # batches a perfect LimitStream pulls: until the running survivor count reaches the limit
                idx = int(np.searchsorted(np.cumsum(keep), limit))
                rec["batches_ideal"] = min(n_batches, idx // batch_rows + 1)
                if rows_out != min(limit, int(keep.sum())):
                    raise SystemExit(f"bench.py: c4 {label}: {rows_out} rows out")
            else:
                if rows_out != int(keep.sum()):
                    raise SystemExit(f"bench.py: c4 {label}: {rows_out} rows out, expected {int(keep.sum())}")
                needed = n * 24   # k, a, b (flag is neither predicate nor projected: it never crosses)
                rec["stream_ms"] = stream_ms          # first push .. last kernel done (the collect() concat + D2H come after)
                rec["input_gbs"] = needed / stream_ms / 1e6
                if transfer == capi.TRANSFER_STAGED:
                    copy_ms = stats["h2d_bytes"] / copy_gbs / 1e6
                    rec["h2d_gbs"] = stats["h2d_bytes"] / stream_ms / 1e6
                    rec["h2d_frac_of_pinned_copy"] = rec["h2d_gbs"] / copy_gbs
                    rec["kernel_ms"] = kern_ms
                    # kernel time hidden under the transfers / total kernel time (1 = perfectly overlapped, 0 = serialised):
                    # serialised, the stream phase would last copy + kernels
                    rec["overlap_ratio"] = max(0.0, min(1.0, (kern_ms + copy_ms - stream_ms) / kern_ms)) if kern_ms > 0 else None
            out["runs"].append(rec)
        del structs
        for x in (kb, ab, bb, fb, *out_pin):
            x.free()
    return out


# ------------------------------------------------------------------------------------------ c1: BASELINE configs[0]
def run_c1(args):
    """1 M-row {name: String, age: Int64}: LazyFrame.from_dataframe(df).filter(age > 25).select([name]).collect() through the host
    layer (logical_plan/builder.rs:96-104 -> physical_plan/plan.rs:97-150 shapes), wall clock including the DataFrame clone, the
    upload, the kernels and the download into a host DataFrame; the oracle's eager engine runs the same call on 1 core."""
    import numpy as np
    from oracle import oracle as O
    from rivulus_b200 import capi
    from rivulus_b200 import frame as F
    n = 1_000_000
    spec = [("name", capi.SYNTH_STR, 0, 0), ("age", capi.SYNTH_AGE100, 1, 0)]

    def q(mod, df):
        return mod.LazyFrame.from_dataframe(df).filter(mod.col("age").gt(mod.lit(25))).select([mod.col("name")]).collect()
    gdf, cdf = F.DataFrame.synth(spec, n), O.DataFrame.synth(spec, n)
    gt, got = [], None
    for r in range(6):
        t0 = time.perf_counter()
        got = q(F, gdf)
        if r > 0:
            gt.append(time.perf_counter() - t0)
    ct, want = [], None
    for r in range(3):
        t0 = time.perf_counter()
        want = q(O, cdf)
        ct.append(time.perf_counter() - t0)
    g, w = got.column_raw(0), want.column_raw(0)
    if (got.height(), got.column_names(), got.dtypes()) != (want.height(), want.column_names(), want.dtypes()) or \
            not all(np.array_equal(x, y) for x, y in zip(g, w)):
        raise SystemExit("bench.py: c1 GPU collect() differs from the oracle's")
    gs, cs = median(gt), median(ct)
    return {"workload": "configs[0]: 1 M rows {name: String, age: Int64}; from_dataframe(df).filter(age > 25).select([name]).collect()",
            "rows": n, "survivors": got.height(), "gpu_wall_ms": gs * 1e3, "gpu_rows_per_s": n / gs,
            "cpu_wall_ms": cs * 1e3, "cpu_rows_per_s": n / cs, "cpu_cores": 1, "cpu_kind": "port (oracle eager engine)",
            "speedup": cs / gs, "parity": "tags, offsets and bytes of the result column == oracle collect()",
            "timing": "host wall clock around the whole collect() call: DataFrame clone, H2D, kernels, D2H into a host DataFrame"}


def run_csv(args):
    """LazyFrame.from_csv(path).filter(flag).select([s, x, id]).collect_streaming() over a 1 M-line file
    {id: Int64, x: Float64, s: String, flag: Boolean} (10 % null fields), default (adaptive) batch size.  Wall clock of the whole
    call — file read (page cache), parse into Arrow buffers, H2D, kernels, concat — against the oracle's restatement of
    execution/file_stream.rs + the streaming operators on 1 core."""
    import tempfile
    import numpy as np
    from oracle import oracle as O
    from rivulus_b200 import frame as F
    n = 1_000_000
    rng = np.random.default_rng(7)
    ids = rng.integers(-10 ** 9, 10 ** 9, n)
    xs = rng.normal(size=n) * 1000.0
    sl = rng.integers(0, 100000, n)
    fl = rng.integers(0, 2, n)
    nul = rng.random((n, 4)) < 0.1
    lines = ["id,x,s,flag"]
    for i in range(n):
        lines.append("%s,%s,%s,%s" % ("" if nul[i, 0] else ids[i], "" if nul[i, 1] else repr(float(xs[i])), "null" if nul[i, 2] else "name_%d" % sl[i],
                                      "" if nul[i, 3] else ("true" if fl[i] else "false")))
    d = tempfile.mkdtemp(prefix="rvl_bench_csv_")
    path = os.path.join(d, "bench.csv")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    size = os.path.getsize(path)
    schema = [("id", F.DT_INT64), ("x", F.DT_FLOAT64), ("s", F.DT_STRING), ("flag", F.DT_BOOLEAN)]

    def q(mod):
        return mod.LazyFrame.from_csv(path, schema).filter(mod.col("flag")).select([mod.col("s"), mod.col("x"), mod.col("id")]).collect_streaming()
    gt, got = [], None
    for r in range(4):
        t0 = time.perf_counter()
        got = q(F)
        if r > 0:
            gt.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    want = q(O)
    cs = time.perf_counter() - t0
    same = got.num_rows() == want.num_rows() and got.column_names() == want.column_names()
    for a, b in zip(got.columns(), want.columns()):
        for x, y in ((a.values, b.values), (a.validity, b.validity), (a.offsets, b.offsets), (a.data, b.data)):
            same = same and ((x is None) == (y is None)) and (x is None or np.array_equal(np.asarray(x), np.asarray(y)))
        same = same and a.null_count == b.null_count
    try:
        os.remove(path); os.rmdir(d)
    except OSError:
        pass
    if not same:
        raise SystemExit("bench.py: csv GPU collect_streaming() differs from the oracle's")
    gs = median(gt)
    return {"workload": "from_csv(1 M lines {id, x, s, flag}, 10 % null fields).filter(flag).select([s, x, id]).collect_streaming()",
            "file_bytes": size, "rows": n, "survivors": got.num_rows(), "gpu_wall_ms": gs * 1e3, "gpu_mb_per_s": size / 1e6 / gs,
            "gpu_rows_per_s": n / gs, "cpu_wall_ms": cs * 1e3, "cpu_mb_per_s": size / 1e6 / cs, "cpu_cores": 1,
            "cpu_kind": "port (oracle CsvFileStream + FilterStream + SelectStream)", "speedup": cs / gs,
            "parity": "values, validity bitmaps, offsets, bytes and null counts of every result column == oracle",
            "note": "the host-side parser (one thread) bounds this path, not the device: see DESIGN.md"}


def run_join(args, ctx):
    """inner join (physical_plan/plan.rs:174-284) through rvl_hash_join_inner, device-resident inputs:
    build side 16 M rows {key: Int64 (a permutation), payload}, probe side 64 M rows {key uniform in [0, 20 M), payload}; output = probe
    payload + build payload per matching pair (80 % of the probe rows).  Checked in full against the numpy model; the oracle's
    HashMap<AnyValue, Vec<usize>> restatement is timed on a 1/80 sample through LazyFrame.inner_join(..).collect() on 1 core."""
    import numpy as np
    from oracle import oracle as O
    from rivulus_b200 import capi
    nb, npr, span = 16_000_000, 64_000_000, 20_000_000
    rng = np.random.default_rng(11)
    bkeys = rng.permutation(nb).astype(np.int64)
    pkeys = rng.integers(0, span, npr).astype(np.int64)
    bpay = rng.integers(-2 ** 62, 2 ** 62, nb).astype(np.int64)
    ppay = np.arange(npr, dtype=np.int64)
    build = ctx.upload([capi.Column(capi.INT64, nb, 0, bkeys), capi.Column(capi.INT64, nb, 0, bpay)])
    probe = ctx.upload([capi.Column(capi.INT64, npr, 0, pkeys), capi.Column(capi.INT64, npr, 0, ppay)])
    ts, out = [], None
    for r in range(4):
        ctx.synchronize()
        t0 = time.perf_counter()
        o = ctx.hash_join_inner(build, 0, probe, 0, [1], [1])
        ctx.synchronize()
        if r > 0:
            ts.append(time.perf_counter() - t0)
        if out is not None:
            out.release()
        out = o
    got = out.download()
    inv = np.full(span, -1, np.int64)
    inv[bkeys] = np.arange(nb, dtype=np.int64)
    hit = inv[pkeys]
    keep = hit >= 0
    same = out.num_rows() == int(keep.sum()) and np.array_equal(got[0].values, ppay[keep]) and np.array_equal(got[1].values, bpay[hit[keep]])
    out.release(); build.release(); probe.release()
    if not same:
        raise SystemExit("bench.py: join result differs from the model")
    # CPU arm on a sample of the same shape
    sb, sp = nb // 80, npr // 80
    l = O.DataFrame.new([("k", [int(x) for x in bkeys[:sb] % sb]), ("bp", [int(x) for x in bpay[:sb]])])
    rdf = O.DataFrame.new([("k", [int(x) for x in pkeys[:sp] % (sb * 5 // 4)]), ("pp", [int(x) for x in ppay[:sp]])])
    t0 = time.perf_counter()
    cj = O.LazyFrame.from_dataframe(l).inner_join(O.LazyFrame.from_dataframe(rdf), "k", "k").collect()
    cs = time.perf_counter() - t0
    gs = median(ts)
    return {"workload": "inner join, build 16 M rows (unique Int64 keys) x probe 64 M rows (80 % match), payload columns gathered from both sides",
            "build_rows": nb, "probe_rows": npr, "pairs": int(keep.sum()), "gpu_wall_ms": gs * 1e3, "gpu_probe_rows_per_s": npr / gs,
            "cpu_sample": f"{sb} x {sp} rows, LazyFrame.inner_join(..).collect() of the oracle (1 core)", "cpu_wall_ms": cs * 1e3,
            "cpu_probe_rows_per_s": sp / cs, "cpu_pairs": cj.height(), "speedup_per_probe_row": (npr / gs) / (sp / cs),
            "parity": "pair count and both gathered columns == numpy model (probe order, build order within a probe row)",
            "timing": "host wall clock around rvl_hash_join_inner with device-resident inputs (sort of the build side, probe, pair fill, gathers)"}


def host_mem_available_bytes():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable:"):
                return int(ln.split()[1]) * 1024
    except Exception:
        pass
    return 64 << 30


def run_e2e(args, ctx, table, preds, proj, rank, world, local_rank, barrier):
    """Same sweep with the table in pinned host memory: every step pushes every batch over PCIe (H2D) through the
    double-buffered streaming executor and reads every result batch back (D2H)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from rivulus_b200 import capi
    import ctypes as C

    rows = args.rows
    # the same host-resident table size per GPU at every N (so the 1 -> 8 curve compares like with like); only a host that cannot
    # pin world x that much shrinks it, and the record says so
    budget = host_mem_available_bytes() * 0.60 / max(world, 1)
    want = min(rows, args.e2e_rows)
    e2e_rows = int(min(want, budget // (BYTES_PER_ROW_IN + 1)))
    rows_reduced = e2e_rows < want
    e2e_rows = max(e2e_rows // 64 * 64, 64)
    batch_rows = min(args.e2e_batch_rows, e2e_rows)
    sub = table.slice(0, e2e_rows)
    host_bufs, host_cols = [], []
    for j in range(5):
        buf = capi.PinnedBuffer(e2e_rows * 8)
        arr = buf.view(np.int64 if j in (0, 1, 3) else np.float64, e2e_rows)
        s = capi.Column(capi.INT64 if j in (0, 1, 3) else capi.FLOAT64, e2e_rows, 0, arr).as_struct()
        capi.check(capi.lib().rvl_batch_download_column(ctx._h, sub._h, j, C.byref(s)))
        host_bufs.append(buf); host_cols.append(arr)
    out_bufs = [capi.PinnedBuffer(batch_rows * 8) for _ in range(4)]
    out_views = [b.view(np.int64, batch_rows) for b in out_bufs]
    dtypes = [capi.INT64, capi.INT64, capi.FLOAT64, capi.INT64, capi.FLOAT64]
    n_batches = (e2e_rows + batch_rows - 1) // batch_rows

    def make_structs(b):
        off = b * batch_rows
        ln = min(batch_rows, e2e_rows - off)
        arr = (capi.RvlColumn * 5)()
        for j in range(5):
            c = capi.Column(dtypes[j], ln, off, host_cols[j]).as_struct()
            arr[j] = c
        return arr
    structs = [make_structs(b) for b in range(n_batches)]
    d2h = 0
    staged = 0
    per_query = []
    transfer = {"auto": capi.TRANSFER_AUTO, "staged": capi.TRANSFER_STAGED, "zero_copy": capi.TRANSFER_ZERO_COPY}[args.e2e_transfer]

    def drain(st, out_rows):
        nonlocal d2h
        b = st.next_batch()
        if b is None:
            return out_rows
        n = b.num_rows()
        for j in range(4):
            s = capi.Column(capi.INT64 if j in (0, 2) else capi.FLOAT64, n, 0, out_views[j] if j in (0, 2) else out_views[j].view(np.float64)).as_struct()
            capi.check(capi.lib().rvl_batch_download_column(ctx._h, b._h, j, C.byref(s)))
        d2h += n * 32
        b.release()
        return out_rows + n

    def e2e_step():
        nonlocal staged
        total = []
        per_query.clear()
        for p in preds:
            tq = time.perf_counter()
            st = ctx.open_stream(dtypes, p, proj, -1, batch_rows=batch_rows, n_staging=3, transfer=transfer)
            out_rows, inflight = 0, 0
            for b in range(n_batches):
                st.push_structs(structs[b], 5)
                inflight += 1
                if inflight >= 2:
                    out_rows = drain(st, out_rows); inflight -= 1
            while inflight > 0:
                out_rows = drain(st, out_rows); inflight -= 1
            staged += st.stats()["h2d_bytes"]
            st.close()
            per_query.append((time.perf_counter() - tq) * 1e3)
            total.append(out_rows)
        return total

    steps = max(1, min(args.steps, args.e2e_steps))
    e2e_step()  # warm-up (allocations, pool growth)
    barrier(); torch.cuda.synchronize()
    d2h = 0
    staged = 0
    t0 = time.perf_counter()
    stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=torch.device("cuda", local_rank))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        totals = e2e_step()
    ctx.synchronize()
    e1.record(stream); e1.synchronize()
    wall = time.perf_counter() - t0
    torch.cuda.synchronize(); barrier()
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    value = e2e_rows * world * len(THRESHOLDS) * steps / wall
    h2d = e2e_rows * BYTES_PER_ROW_IN * len(THRESHOLDS)
    for b in host_bufs + out_bufs:
        b.free()
    in_place = staged // steps < h2d
    return {"value": value, "unit": "rows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h // steps, "steps": steps,
            "rows_per_gpu": e2e_rows, "rows_reduced_by_host_ram": rows_reduced, "batch_rows": batch_rows, "ms_per_step": wall * 1000.0 / steps,
            "h2d_gbs": h2d * steps / wall / 1e9, "survivors": totals, "transfer": args.e2e_transfer, "per_query_ms_last_step": [round(x, 2) for x in per_query],
            "h2d_copy_engine_bytes_per_step": staged // steps,
            "h2d_note": ("h2d_bytes_per_step = the host-resident input of the step (every column of every query); the predicate column "
                         "crosses on the copy engine (h2d_copy_engine_bytes_per_step), the projected columns are read in place by the "
                         "kernels over PCIe, which fetch only the lines holding survivors (dense tiles whole)") if in_place else
                        "every input byte crosses on the copy engine",
            "path": "rvl_stream_open/push/next + rvl_batch_download_column, pinned host buffers, 3 staging slots",
            "timing": "host wall clock around the synchronised region (H2D, kernels, D2H all inside)"}


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print to fd 1 (NCCL's version banner, ...) goes to stderr; emit() writes the ONE JSON line to the real stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows per GPU (BASELINE configs[1]: 1e9)")
    ap.add_argument("--cpu-rows", type=int, default=None, help="rows of the bounded CPU sample (per shard)")
    ap.add_argument("--verify-rows", type=int, default=16_000_000)
    ap.add_argument("--e2e-rows", type=int, default=512_000_000, help="rows per GPU of the host-resident table of the e2e leg (the same at every N)")
    ap.add_argument("--no-configs", action="store_true", help="skip the c1/c3/c4/c5 sub-records")
    ap.add_argument("--no-golden", action="store_true", help="skip the full-size golden checksum comparison")
    ap.add_argument("--trace", action="store_true", help="stderr: host time and pool size around every operator call of the timed steps")
    ap.add_argument("--c5-rows", type=int, default=C5_ROWS)
    ap.add_argument("--c5-reps", type=int, default=3)
    ap.add_argument("--c3-rows", type=int, default=200_000_000)
    ap.add_argument("--c3-reps", type=int, default=3)
    ap.add_argument("--c3-verify-rows", type=int, default=4_000_000)
    ap.add_argument("--c4-batches", type=int, default=64)
    ap.add_argument("--e2e-batch-rows", type=int, default=16 << 20)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-transfer", default="auto", choices=["auto", "staged", "zero_copy"], help="rvl_stream_config.transfer of the e2e leg")
    ap.add_argument("--plan", default="auto", choices=["auto", "fused", "two_pass"], help="execution plan of rvl_filter_project (rvl_plan)")
    ap.add_argument("--sparse-max", type=int, default=None)
    ap.add_argument("--dense-slots", type=int, default=None)
    ap.add_argument("--dense-ctas", type=int, default=None)
    ap.add_argument("--scan-warps", type=int, default=None)
    ap.add_argument("--dense-warps", type=int, default=None)
    ap.add_argument("--scan-slots", type=int, default=None)
    args = ap.parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # `python bench.py --gpus N` without a launcher: start one rank per GPU ourselves, exactly as the driver would
        import subprocess
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    quiet_stdout()
    if args.cpu_rows is None:
        args.cpu_rows = 500_000 if args.impl == "reference" else 4_000_000
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
