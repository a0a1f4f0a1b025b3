set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-e2e --cpu-rows 0 --verify-rows 0 > gpurun_out/b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --cpu-rows 0 --verify-rows 0 > gpurun_out/ncu_launches.log 2>&1
wc -l gpurun_out/bench_full.json gpurun_out/bench_ref.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_full.json"))
print(d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]/1e9, d["e2e"]["per_query_ms_last_step"], d["gpu_launches"], d["clocks"])
PY
