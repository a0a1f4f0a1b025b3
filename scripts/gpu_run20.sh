python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu17.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu17.log
timeout 600 python scripts/bench_configs.py --which c3,c5 > gpurun_out/configs_r01b.json 2> gpurun_out/configs_r01b.err; echo "rc=$?"
tail -3 gpurun_out/configs_r01b.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/configs_r01b.json"))
for k in ("c3","c5"):
    for q in d.get(k,[]): print(k, q["label"], round(q["device_ms"],3), "ms", round(q["alg_gbs"]), "GB/s", round(q["frac_of_peak"],3))
PY
python bench.py --steps 5 --warmup 3 --no-e2e --cpu-rows 0 --verify-rows 0 > gpurun_out/b_tmp.json 2> gpurun_out/b_tmp.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/b_tmp.json"))
print("ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4), [ (round(x["kernel_ms"],3), round(x["frac_of_peak"],3)) for x in d["sweep"]])
PY
