python scripts/join_one.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_join_launches.csv python scripts/join_one.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_join_launches.csv')) if len(r)>10 and r[0].isdigit()]
hdr=[r for r in csv.reader(open('gpurun_out/r02_join_launches.csv')) if r and r[0]=="ID"][0]
iv,iu,ik=hdr.index("Metric Value"),hdr.index("Metric Unit"),hdr.index("Kernel Name")
half=rows[len(rows)//2:]   # the second (warm) invocation
agg=collections.OrderedDict()
for r in half:
    ms=float(r[iv].replace(",",""))*{"ns":1e-6,"us":1e-3,"ms":1,"s":1e3}.get(r[iu],1e-6)
    k=r[ik].split("(")[0].replace("void ","")[:70]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=ms
tot=sum(a[1] for a in agg.values())
print("total device ms of one join:", round(tot,3))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print(f"{a[1]:8.3f} ms  {100*a[1]/tot:5.1f} %  x{a[0]:<3d} {k}")
PY
