"""Where the time of LazyFrame.from_csv(..).filter(flag).select(..).collect_streaming() goes (1 M lines, 42 MB)."""
import os, sys, time, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rivulus_b200 import frame as F
n = 1_000_000
rng = np.random.default_rng(7)
ids = rng.integers(-10 ** 9, 10 ** 9, n); xs = rng.normal(size=n) * 1000.0; sl = rng.integers(0, 100000, n); fl = rng.integers(0, 2, n); nul = rng.random((n, 4)) < 0.1
lines = ["id,x,s,flag"]
for i in range(n):
    lines.append("%s,%s,%s,%s" % ("" if nul[i, 0] else ids[i], "" if nul[i, 1] else repr(float(xs[i])), "null" if nul[i, 2] else "name_%d" % sl[i], "" if nul[i, 3] else ("true" if fl[i] else "false")))
path = os.path.join(tempfile.mkdtemp(), "b.csv")
open(path, "w").write("\n".join(lines) + "\n")
schema = [("id", F.DT_INT64), ("x", F.DT_FLOAT64), ("s", F.DT_STRING), ("flag", F.DT_BOOLEAN)]
exe = "/tmp/csv_parse_speed"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
subprocess.check_call(["g++", "-O2", "-std=c++17", "-I" + root + "/rivulus_b200/host", "-I" + root + "/include", root + "/scripts/csv_parse_speed.cpp", "-o", exe,
                       "-L" + root + "/rivulus_b200/lib", "-lrivulus_host", "-lrivulus_gpu", "-Wl,-rpath," + root + "/rivulus_b200/lib", "-pthread"])
for t in (0, 1, 2, 4, 8, 16):
    print("parser alone, threads", t, subprocess.run([exe, path, str(t)], capture_output=True, text=True).stdout.strip().splitlines()[-1], flush=True)
def q(sel, fusion=True, threads=-1, batch=None):
    F.set_stream_fusion(fusion); F.set_csv_threads(threads)
    ts = []
    for r in range(4):
        t0 = time.perf_counter()
        lf = F.LazyFrame.from_csv(path, schema, batch).filter(F.col("flag")).select([F.col(c) for c in sel])
        out = lf.collect_streaming(); rows = out.num_rows()
        ts.append(time.perf_counter() - t0)
    F.set_stream_fusion(True); F.set_csv_threads(-1)
    return rows, [round(x * 1e3, 1) for x in ts]
os.environ["RVL_HOST_TRACE"] = "1"
print("fused  [s,x,id] default threads", q(["s", "x", "id"]), flush=True)
print("fused  [s,x,id] 0 threads      ", q(["s", "x", "id"], threads=0), flush=True)
print("fused  [x,id]   (no string out)", q(["x", "id"]), flush=True)
print("chain  [s,x,id]                ", q(["s", "x", "id"], fusion=False), flush=True)
print("fused  [s,x,id] batch 1000000  ", q(["s", "x", "id"], batch=1000000), flush=True)
print("fused  [s,x,id] batch 10000    ", q(["s", "x", "id"], batch=10000), flush=True)
