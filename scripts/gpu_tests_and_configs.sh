python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu18.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu18.log
timeout 600 python scripts/bench_configs.py --which c3,c5 > gpurun_out/configs_r01c.json 2> gpurun_out/configs_r01c.err; echo "rc=$?"
tail -3 gpurun_out/configs_r01c.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/configs_r01c.json"))
for k in ("c3","c5"):
    for q in d.get(k,[]): print(k, q["label"], round(q["device_ms"],3), "ms", round(q["alg_gbs"]), "GB/s", round(q["frac_of_peak"],3))
PY
