set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"])
for q in d["sweep"]: print(q["threshold"], [round(x,3) for x in q["kernel_ms_min_median_max"]], round(q["frac_of_peak"],3))
print("e2e", d["e2e"]["value"]/1e9)
print("csv", {k:v for k,v in d["csv"].items() if k in ("gpu_wall_ms","gpu_mb_per_s","cpu_wall_ms","speedup","error")})
print("join", json.dumps(d.get("join"))[:900])
PY
python -c "import __graft_entry__ as g; g.smoke()"
