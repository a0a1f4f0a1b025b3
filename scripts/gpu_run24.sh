python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu19.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu19.log
for t in staged zero_copy; do
python bench.py --steps 3 --warmup 3 --cpu-rows 0 --verify-rows 0 --e2e-transfer $t > gpurun_out/b_e2e_$t.json 2> gpurun_out/b_e2e_$t.err
python - <<PY
import json
d=json.load(open("gpurun_out/b_e2e_$t.json"))
e=d["e2e"]
print("$t", "e2e rows/s", round(e["value"]/1e9,3), "G  ms/step", round(e["ms_per_step"],1), "CE bytes", e.get("h2d_copy_engine_bytes_per_step"), "frac", round(d["roofline"]["frac"],3))
PY
done
for t in staged zero_copy; do
timeout 600 python scripts/bench_configs.py --which c4 --transfer $t > gpurun_out/configs_c4_$t.json 2> gpurun_out/configs_c4_$t.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/configs_c4_$t.json"))
for q in d.get("c4",[]): print("c4 $t", q["batch_rows"], q["query"], round(q["wall_ms"],3), "ms", q["batches_transferred"], "/", q["batches_ideal"], round(q["stream_bytes_total"]/q["wall_ms"]/1e6,1), "GB/s of input")
PY
done
