python scripts/profile_one.py --rows 1000000000 --plan two_pass > gpurun_out/plain2.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_tp.csv python scripts/profile_one.py --rows 1000000000 --plan two_pass > gpurun_out/ncu_tp.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/launches_tp.csv")))
i=[k for k,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[i]
for r in rows[i+1:]:
    d=dict(zip(h,r))
    if "gen_" in d["Kernel Name"]: continue
    print(d["ID"], d["Kernel Name"][:50], d["Metric Name"], d["Metric Value"], d["Metric Unit"])
PY
