"""N concurrent scripts/microbench_duplex instances, one per GPU, optionally each bound to its GPU's CPU set (NVML affinity):
the aggregate H2D / D2H rate the HOST sustains when 1 / 2 / 4 / 8 GPUs stream at once — the ceiling of bench.py's e2e leg.

    python scripts/pcie_scaling.py --gpus 8 [--affinity] > profiles/r02_pcie_scaling_n8.txt
"""
import argparse
import os
import re
import subprocess
import sys
import time

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--affinity", action="store_true")
ap.add_argument("--window", type=float, default=1.5)
args = ap.parse_args()
here = os.path.dirname(os.path.abspath(__file__))
exe = os.path.join(here, "microbench_duplex")
cpusets = [None] * args.gpus
if args.affinity:
    import pynvml
    pynvml.nvmlInit()
    ncpu = os.cpu_count() or 1
    for g in range(args.gpus):
        h = pynvml.nvmlDeviceGetHandleByIndex(g)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1 and 64 * i + b < ncpu]
        cpusets[g] = cpus or None
start = time.time() + 6.0 + 1.0 * args.gpus    # pinning 4 GiB per process takes a few seconds
procs = []
for g in range(args.gpus):
    def pre(cpus=cpusets[g]):
        if cpus:
            os.sched_setaffinity(0, cpus)
    procs.append(subprocess.Popen([exe, str(g), f"{start:.3f}", str(args.window)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, preexec_fn=pre))
rows = {}
for g, p in enumerate(procs):
    out, err = p.communicate()
    if err.strip():
        print(err.strip(), file=sys.stderr)
    for ln in out.splitlines():
        m = re.match(r"dev (\d+) window (\d+) (.*?)\s+h2d\s+([\d.]+) GB/s\s+d2h\s+([\d.]+) GB/s", ln)
        if m:
            rows.setdefault(int(m.group(2)), []).append((m.group(3).strip(), float(m.group(4)), float(m.group(5))))
print(f"# {args.gpus} GPU(s) streaming at once, affinity={'GPU cpu set: ' + str([len(c) if c else None for c in cpusets]) if args.affinity else 'unbound'}, host cpus={os.cpu_count()}")
print("window                               per-GPU h2d (min/mean)   per-GPU d2h (min/mean)   aggregate h2d   aggregate d2h   aggregate both")
for w in sorted(rows):
    name = rows[w][0][0]
    h = [r[1] for r in rows[w]]; d = [r[2] for r in rows[w]]
    print(f"{name:36s} {min(h):7.2f} / {sum(h)/len(h):7.2f}        {min(d):7.2f} / {sum(d)/len(d):7.2f}        {sum(h):8.2f}        {sum(d):8.2f}        {sum(h)+sum(d):8.2f}")
