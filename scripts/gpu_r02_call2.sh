set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_synth_checksums.py tests/test_host_golden.py -m gpu -x -q -k "string or str or config3 or concat or golden or stream or take or overflow or two_pass" > gpurun_out/r02_pytest_strings.log 2>&1; echo "pytest-strings rc=$?"; tail -15 gpurun_out/r02_pytest_strings.log
# sparse/dense crossover with the 64-byte-granule gather
for sm in 224 320 416 512 640; do for thr in 899 849 799 749 699; do echo "sparse_max=$sm thr=$thr"; python scripts/profile_one.py --rows 1000000000 --thresholds $thr --reps 3 --sparse-max $sm | tail -1; done; done > gpurun_out/r02_sparse_sweep.txt 2>&1
# string kernels: round-1 pair vs round-2 pair, dense_min sweep
python - > gpurun_out/r02_string_ab.txt 2>&1 <<'PY'
import sys, json
sys.path.insert(0, '.')
import bench
from rivulus_b200 import capi
class A: c3_rows=200_000_000; c3_reps=3; c3_verify_rows=2_000_000
ctx = capi.Context(0)
for kern, dm in ((1, 128), (2, 0), (2, 64), (2, 128), (2, 256), (2, 1025)):
    ctx.set_option(capi.OPT_STRING_KERNEL, kern); ctx.set_option(capi.OPT_STRING_DENSE_MIN, dm)
    r = bench.run_c3(A, ctx, 6535.7)
    print(kern, dm, [(q['label'], round(q['device_ms'], 3), round(q['frac'], 3)) for q in r['queries']], r['parity'][:20], flush=True)
PY
cat gpurun_out/r02_string_ab.txt
timeout 900 python bench.py > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1_b.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_n1_b.json"))
print(d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], [round(q["frac_of_peak"],3) for q in d["sweep"]], [round(q["kernel_ms"],3) for q in d["sweep"]])
print("e2e", d["e2e"]["value"]/1e9)
for q in d["c5"]["queries"]: print("c5", q["label"], q["device_ms"], q["frac"], q["gather"]["ms"], q["gather"]["gbs"])
for q in d["c3"]["queries"]: print("c3", q["label"], q["device_ms"], q["frac"])
for r in d["c4"]["runs"]: print("c4", r["batch_rows"], r["query"], r["wall_ms"], r.get("stream_ms"), r["batches_transferred"], r.get("batches_ideal"), r["operator_launches"], r.get("h2d_gbs"), r.get("input_gbs"), r.get("overlap_ratio"))
print("c1", d["c1"])
PY
