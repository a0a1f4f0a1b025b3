"""Which host-side action makes the first operator call afterwards slow?  (bench.py saw one 3-540 ms call at the start of its timed region.)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rivulus_b200 import capi
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_I64, 1, 0), (capi.SYNTH_F64, 2, 0), (capi.SYNTH_I64, 3, 0), (capi.SYNTH_F64, 4, 0)]
rows = 1_000_000_000
t = ctx.gen_batch(spec, rows, 0)
preds = [capi.predicate(0, ">", thr) for thr in (998, 899, 499, 99)]
def step():
    out_ms = []
    for p in preds:
        t0 = time.perf_counter()
        o = ctx.filter_project(t, p, [1, 2, 3, 4]); o.num_rows()
        out_ms.append(round((time.perf_counter() - t0) * 1e3, 2)); o.release()
    return out_ms
def small():
    sl = t.slice(0, 16_000_000)
    for thr in (998, 499):
        o = ctx.filter_project(sl, capi.predicate(0, ">", thr), [1, 2, 3, 4]); o.num_rows(); [o.checksum(j) for j in range(4)]; o.release()
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode != "nosmall": small()
for _ in range(3): step()
print("steady", step(), ctx.pool_stats()[0] >> 20, flush=True)
ctx.synchronize(); print("after stream sync", step(), flush=True)
ctx.profile_enable(True); print("profile on", step(), flush=True); ctx.profile_read_launches(); print("after profile read", step(), flush=True)
torch.cuda.synchronize(); print("after torch.cuda.synchronize", step(), flush=True)
print("again", step(), flush=True)
torch.cuda.synchronize(); print("after 2nd torch.cuda.synchronize", step(), flush=True)
s = torch.cuda.ExternalStream(ctx.cuda_stream(), device=torch.device("cuda", 0)); e = torch.cuda.Event(enable_timing=True); e.record(s)
print("after event record", step(), flush=True)
time.sleep(0.3); print("after sleep 0.3", step(), flush=True)
import pynvml; pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetPowerUsage(h)
print("after nvml", step(), flush=True)
