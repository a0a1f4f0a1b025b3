timeout 600 python scripts/bench_configs.py --which c3,c5,c4 > gpurun_out/configs_r01a.json 2> gpurun_out/configs_r01a.err; echo "rc=$?"
tail -3 gpurun_out/configs_r01a.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/configs_r01a.json"))
for k in ("c3","c5"):
    for q in d.get(k,[]): print(k, q["label"], round(q["device_ms"],3), "ms", round(q["alg_gbs"]), "GB/s", round(q["frac_of_peak"],3))
for q in d.get("c4",[]): print("c4", q["batch_rows"], q["query"], round(q["wall_ms"],3), "ms", q["batches_transferred"], "/", q["batches_ideal"], round(q["h2d_gbs"],1), "GB/s")
PY
