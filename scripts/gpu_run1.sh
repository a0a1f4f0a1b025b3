set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?" 
tail -15 gpurun_out/pytest_gpu8.log
for plan in fused two_pass; do
  python bench.py --steps 5 --warmup 3 --no-e2e --cpu-rows 0 --verify-rows 16000000 --plan $plan > gpurun_out/bench_$plan.json 2> gpurun_out/bench_$plan.err; echo "rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$plan.json"))
print("$plan", d["ms_per_step"], d["roofline"]["frac"], [ (round(x["kernel_ms"],3), round(x["frac_of_peak"],3)) for x in d["sweep"]])
PY
done
