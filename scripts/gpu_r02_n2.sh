set -x
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sharded or gather" > gpurun_out/r02_pytest_n2.log 2>&1; echo "pytest-n2 rc=$?"; tail -5 gpurun_out/r02_pytest_n2.log
timeout 900 python bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n2.json"))
print(d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], [round(q["frac_of_peak"],3) for q in d["sweep"]])
print("e2e", d["e2e"]["value"]/1e9, d["e2e"]["rows_per_gpu"], d["e2e"].get("h2d_gbs"))
print("cpu", d.get("cpu_baseline"))
print(json.dumps(d["c5"], indent=0)[:3000])
PY
