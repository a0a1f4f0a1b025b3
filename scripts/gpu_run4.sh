summ() { python - "$1" <<'PY'
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
i=[k for k,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[i]
for r in rows[i+1:]:
    d=dict(zip(h,r))
    if "compact_dense" in d["Kernel Name"] and d["Metric Name"]=="gpu__time_duration.sum": print(d["ID"], d["Kernel Name"][:30], d["Metric Value"])
PY
}
for v in 0 1; do
  if [ $v = 1 ]; then export RVL_DEBUG_SORT_LISTS=1; fi
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_sort$v.csv python scripts/profile_one.py --rows 1000000000 --plan two_pass --thresholds 899,499,99 > gpurun_out/ncu_sort$v.log 2>&1
  echo "sorted=$v"; summ gpurun_out/l_sort$v.csv
done
