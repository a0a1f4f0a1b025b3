python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu15.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu15.log
python bench.py --steps 10 --warmup 3 --no-e2e --cpu-rows 0 --verify-rows 0 > gpurun_out/b_tmp.json 2> gpurun_out/b_tmp.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/b_tmp.json"))
print("ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4), [ (round(x["kernel_ms"],3), round(x["frac_of_peak"],3)) for x in d["sweep"]])
PY
