python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu12.log
run() { python bench.py --steps 5 --warmup 3 --no-e2e --cpu-rows 0 --verify-rows 0 "$@" > gpurun_out/b_tmp.json 2> gpurun_out/b_tmp.err || tail -3 gpurun_out/b_tmp.err
  python - "$*" <<'PY'
import json,sys
d=json.load(open("gpurun_out/b_tmp.json"))
print(sys.argv[1], "| ms/step", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4), [ (round(x["kernel_ms"],3), round(x["frac_of_peak"],3)) for x in d["sweep"]])
PY
}
run --dense-warps 8
run --dense-warps 16
run --dense-warps 16 --dense-slots 14
