set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu2.log
for k in 1 3 4; do python scripts/c3_one.py --kernel $k --reps 3 | tail -4; done
python - <<'PY'
import sys
sys.path.insert(0, '.')
from rivulus_b200 import capi
ctx = capi.Context(0)
spec = [(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)]
t = ctx.gen_batch(spec, 500_000_000, 3_500_000_000)
ctx.profile_enable(True)
for plan, name in ((capi.PLAN_TWO_PASS, "two_pass"), (capi.PLAN_FUSED, "fused")):
    ctx.set_option(capi.OPT_PLAN, plan)
    for thr in (899, 499):
        for r in range(3):
            o = ctx.filter_project(t, capi.predicate(0, ">", thr), [0, 1, 2])
            ms = ctx.profile_read_launches()
            n = o.num_rows(); o.release()
        print("c5 shard 500M", name, thr, n, "device ms", [round(x, 3) for x in ms], flush=True)
PY
timeout 900 python bench.py > gpurun_out/r02_bench_n1_d.json 2> gpurun_out/r02_bench_n1_d.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1_d.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1_d.json"))
print(d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], [round(q["frac_of_peak"],3) for q in d["sweep"]], [[round(x,3) for x in q["kernel_ms_min_median_max"]] for q in d["sweep"]])
print("e2e", d["e2e"]["value"]/1e9)
for q in d["c5"]["queries"]: print("c5", q["label"], q["device_ms"], q["frac"], q["gather"]["ms"], q["gather"]["gbs"])
for q in d["c3"]["queries"]: print("c3", q["label"], q["device_ms"], q["frac"])
print("c1", d["c1"]["gpu_wall_ms"], d["c1"]["speedup"])
PY
