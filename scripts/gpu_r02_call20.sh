RVL_TRACE_ALLOC=1 python scripts/outlier_probe.py all 2>&1 | tail -14
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"])
for q in d["sweep"]: print(q["threshold"], q["kernel_ms_min_median_max"])
print("e2e", d["e2e"]["value"]/1e9)
for k in ("c3","c5"): print(k, [(q["label"], round(q["device_ms"],3), round(q["frac"],3)) for q in d[k]["queries"]])
print([(r["batch_rows"], r["query"][:40], round(r["wall_ms"],3), r.get("input_gbs")) for r in d["c4"]["runs"]])
PY
