set -x
python scripts/profile_one.py --rows 1000000000 --thresholds 998 --reps 3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_998.csv python scripts/profile_one.py --rows 1000000000 --thresholds 998 --reps 2 > /dev/null 2>&1
for k in 1 2; do
python scripts/c3_one.py --kernel $k --reps 2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c3_k$k.csv python scripts/c3_one.py --kernel $k --reps 2 > /dev/null 2>&1
done
python - <<'PY'
import csv, collections
for f in ("r02_launches_998", "r02_launches_c3_k1", "r02_launches_c3_k2"):
    rows = list(csv.reader(open(f"gpurun_out/{f}.csv")))
    hdr = None; out = []
    for r in rows:
        if r and r[0] == "ID": hdr = r; continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r)); out.append((d["Kernel Name"][:60], float(d["Metric Value"].replace(",", "")) / 1e3))
    print("==", f)
    for name, us in out[-24:]: print(f"{us:10.1f} us  {name}")
PY
