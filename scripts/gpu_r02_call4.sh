set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synth_checksums.py tests/test_host_golden.py -m gpu -x -q -k "string or str or config3 or concat or golden or stream or take or overflow or two_pass" > gpurun_out/r02_pytest_strings2.log 2>&1; echo "pytest-strings rc=$?"; tail -15 gpurun_out/r02_pytest_strings2.log
for k in 1 2; do python scripts/c3_one.py --kernel $k --reps 3; done
python scripts/c3_one.py --kernel 2 --dense-min 0 --reps 2
python scripts/c3_one.py --kernel 2 --dense-min 1025 --reps 2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3_k2b.csv python scripts/c3_one.py --kernel 2 --reps 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r02_launches_c3_k2b.csv")))
hdr = None; out = []
for r in rows:
    if r and r[0] == "ID": hdr = r; continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r)); out.append((d["Kernel Name"][:60], float(d["Metric Value"].replace(",", "")) / 1e3))
for name, us in out[-16:]: print(f"{us:10.1f} us  {name}")
PY
timeout 900 python bench.py > gpurun_out/r02_bench_n1_c.json 2> gpurun_out/r02_bench_n1_c.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1_c.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1_c.json"))
print(d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], [round(q["frac_of_peak"],3) for q in d["sweep"]], [[round(x,3) for x in q["kernel_ms_min_median_max"]] for q in d["sweep"]])
print("e2e", d["e2e"]["value"]/1e9)
for q in d["c5"]["queries"]: print("c5", q["label"], q["device_ms"], q["frac"], q["gather"]["ms"], q["gather"]["gbs"])
for q in d["c3"]["queries"]: print("c3", q["label"], q["device_ms"], q["frac"])
for r in d["c4"]["runs"]:
    if "limit(1000)" in r["query"]: print("c4", r["batch_rows"], r["query"], r["wall_ms"], r["batches_transferred"], r.get("batches_ideal"), r["operator_launches"])
print("c1", d["c1"]["gpu_wall_ms"], d["c1"]["speedup"])
PY
