python scripts/bench_configs.py --which c3,c5 > gpurun_out/configs_final.json 2> gpurun_out/configs_final.err; echo rc=$?
python scripts/bench_configs.py --which c3 --c3-rows 50000000 --reps 0 > gpurun_out/c3_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k 'regex:string_gather|string_sizes' -c 6 -o gpurun_out/prof_r01_str -f python scripts/bench_configs.py --which c3 --c3-rows 50000000 --reps 0 > gpurun_out/ncu_str.log 2>&1
python scripts/ncu_top.py gpurun_out/prof_r01_str.ncu-rep 10 > gpurun_out/prof_r01_str.txt 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/configs_final.json"))
for k in ("c3","c5"):
    for q in d.get(k,[]): print(k, q["label"], round(q["device_ms"],3), "ms", round(q["alg_gbs"]), "GB/s", round(q["frac_of_peak"],3))
PY
