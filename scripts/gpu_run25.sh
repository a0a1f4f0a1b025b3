python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu20.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu20.log
for t in auto zero_copy staged; do
python bench.py --steps 3 --warmup 3 --cpu-rows 0 --verify-rows 0 --e2e-transfer $t > gpurun_out/b_e2e_$t.json 2> gpurun_out/b_e2e_$t.err
python - <<PY
import json
d=json.load(open("gpurun_out/b_e2e_$t.json"))
e=d["e2e"]
print("$t", "e2e rows/s", round(e["value"]/1e9,3), "G  ms/step", round(e["ms_per_step"],1), "CE bytes", e.get("h2d_copy_engine_bytes_per_step"), e.get("per_query_ms_last_step"))
PY
done
