ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_c3.csv python scripts/bench_configs.py --which c3 --c3-rows 50000000 --reps 0 > gpurun_out/ncu_c3.log 2>&1; echo rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_c5.csv python scripts/bench_configs.py --which c5 --c5-rows 500000000 --reps 0 > gpurun_out/ncu_c5.log 2>&1; echo rc=$?
python - <<'PY'
import csv
for f in ("gpurun_out/l_c3.csv","gpurun_out/l_c5.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    print(f)
    for r in rows:
        name=r[4][:60]
        if "gen_" in name or "synth" in name: continue
        print("  ", name, r[-1], r[-2])
PY
