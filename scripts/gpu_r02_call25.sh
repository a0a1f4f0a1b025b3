set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"])
for q in d["sweep"]: print(q["threshold"], q["kernel_ms_min_median_max"], round(q["frac_of_peak"],3))
print("e2e", d["e2e"]["value"]/1e9)
for k in ("c3","c5"): print(k, [(q["label"], round(q["device_ms"],3), round(q["frac"],3)) for q in d[k]["queries"]])
print("csv", json.dumps(d.get("csv"))[:1200])
print("c1", json.dumps(d.get("c1"))[:300])
PY
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"; cat gpurun_out/r02_bench_reference.json | cut -c1-600
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1500 bash scripts/gpu_r02_profile.sh > gpurun_out/r02_profile.log 2>&1; echo "profile rc=$?"; tail -3 gpurun_out/r02_profile.log
