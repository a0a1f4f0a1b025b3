"""profiles/traffic.json from an `ncu --set full` capture of scripts/profile_one.py (one operator invocation per threshold, in order):
dram__bytes_read.sum + dram__bytes_write.sum per kernel, summed per invocation; `dram_bytes_per_launch` (what bench.py reports as
roofline.traffic, labelled static) is the mean over the four invocations of the sweep."""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
thresholds = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "998,899,499,99").split(",")]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
iname, ir, iw, it = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
units = rows[1]


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_ms(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)


kern = []
for r in rows[2:]:
    kern.append({"kernel": r[iname][:48], "ms_under_ncu": round(to_ms(r[it], units[it]), 4), "read": to_bytes(r[ir], units[ir]), "write": to_bytes(r[iw], units[iw])})
# an invocation starts at every predicate_scan / chunk_filter kernel
inv = []
for k in kern:
    if k["kernel"].startswith("void rvl::predicate_scan") or "predicate_scan_kernel" in k["kernel"] or "chunk_filter_kernel" in k["kernel"] or "sample_selectivity" in k["kernel"]:
        if "sample_selectivity" in k["kernel"] or not inv or inv[-1]["closed"]:
            inv.append({"kernels": [], "closed": False})
        inv[-1]["closed"] = "sample_selectivity" not in k["kernel"]
    if inv:
        inv[-1]["kernels"].append(k)
inv = inv[:len(thresholds)]
per = {}
for thr, v in zip(thresholds, inv):
    per[str(thr)] = {"dram_bytes_read": sum(k["read"] for k in v["kernels"]), "dram_bytes_write": sum(k["write"] for k in v["kernels"]), "kernels": v["kernels"]}
mean = sum(p["dram_bytes_read"] + p["dram_bytes_write"] for p in per.values()) / max(len(per), 1)
json.dump({"source": "ncu --set full --clock-control none, scripts/profile_one.py --rows 1000000000 (AUTO plan: predicate_scan + compact_dense + gather_sparse with "
                     "64-byte-granule loads, sparse_max 384), round 2 kernels; per operator invocation",
           "per_threshold": per, "dram_bytes_per_launch": mean}, open(out, "w"), indent=1)
print("dram_bytes_per_launch", mean / 1e9, "GB;", {t: round((p["dram_bytes_read"] + p["dram_bytes_write"]) / 1e9, 2) for t, p in per.items()})
