set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()"
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"])
for q in d["sweep"]: print(q["threshold"], [round(x,3) for x in q["kernel_ms_min_median_max"]], round(q["frac_of_peak"],3))
print("e2e", d["e2e"]["value"]/1e9)
for k in ("c3","c5"): print(k, [(q["label"], round(q["device_ms"],3), round(q["frac"],3)) for q in d[k]["queries"]])
print("c1", d["c1"]["gpu_wall_ms"], d["c1"]["speedup"])
print("csv", {k:v for k,v in d["csv"].items() if k in ("gpu_wall_ms","gpu_mb_per_s","cpu_wall_ms","speedup","error")})
print("join", {k:v for k,v in d["join"].items() if k in ("gpu_wall_ms","gpu_probe_rows_per_s","speedup_per_probe_row","error")})
print([(r["batch_rows"], r["query"][:44], round(r["wall_ms"],3), r.get("input_gbs"), r.get("batches_transferred")) for r in d["c4"]["runs"]])
PY
