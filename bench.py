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
    """Samples SM clock and throttle reasons of one GPU while the timed region runs (pynvml)."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, uuid=None, index=0, period=0.05):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
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
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

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

    def step():
        counts = []
        for p in preds:
            out = ctx.filter_project(table, p, proj)   # kernels + count readback (AUTO plan: scan + compaction passes at this size)
            counts.append(out.num_rows())
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
    barrier(); torch.cuda.synchronize()
    ctx.profile_enable(True)
    launches0 = ctx.launch_count()
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        counts = step()
    e1.record(stream)
    e1.synchronize()
    clocks = sampler.stop()
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
        sweep.append({"threshold": thr, "selectivity": s_act, "survivors": counts[qi], "kernel_ms": avg_ms, "b_alg_gb": alg / 1e9,
                      "alg_gbs": alg / 1e9 / (avg_ms / 1e3) if avg_ms > 0 else None,
                      "frac_of_peak": alg / 1e9 / (avg_ms / 1e3) / peak if avg_ms > 0 else None,
                      "rows_per_s": rows / (avg_ms / 1e3) if avg_ms > 0 else None,
                      "b_scan_gb": (BYTES_PER_ROW_IN * rows + N_PROJ * 8 * counts[qi]) / 1e9})
    achieved = alg_total / 1e9 / (kern_total / 1e3) if kern_total > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
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
                     "peak_source": peak_src,
                     "kernel": ("fused_filter_project_kernel<kPredI64>" if args.plan == "fused" else
                                "predicate_scan_kernel<kPredI64> + compact_dense_kernel + gather_sparse_kernel (one operator invocation)"),
                     "definition": "sum of algorithmic bytes (SURVEY 8(d): 8.16/22.2/54.0/68.8 B per row at 0.1/10/50/90 %, x rows) of the timed "
                                   "operator invocations / sum of their device durations (CUDA events on the library stream around the kernels "
                                   "of each invocation)"},
        "sweep": sweep, "gpu_launches": int(total_launches.item()), "clocks": clocks, "parity": parity,
    }

    # ---- CPU baseline (rank 0, N=1 only): the oracle's eager engine on a bounded sample, 1 core like the reference
    if rank == 0 and world == 1 and args.cpu_rows > 0:
        t0 = time.time()
        v, secs = cpu_eager_sweep(args.cpu_rows, 1, 1, 0)
        line["cpu_baseline"] = {"value": v, "unit": "rows/s", "cores": 1, "kind": "port",
                                "sample": f"{args.cpu_rows} rows x 4 queries, eager collect() restatement ({secs:.1f} s)"}

    # ---- end to end through the streaming C ABI with HOST buffers
    if not args.no_e2e:
        line["e2e"] = run_e2e(args, ctx, table, preds, proj, rank, world, local_rank, barrier)

    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


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
    budget = host_mem_available_bytes() * 0.30 / max(world, 1)
    e2e_rows = args.e2e_rows if args.e2e_rows > 0 else int(min(rows, budget // (BYTES_PER_ROW_IN + 32 * 0.9)))
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
            "rows_per_gpu": e2e_rows, "batch_rows": batch_rows, "ms_per_step": wall * 1000.0 / steps,
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
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows of the host-resident table (default: as many of --rows as fit in 30%% of host RAM)")
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
