python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream" 2>&1 | tail -2
for t in auto; do
timeout 600 python scripts/bench_configs.py --which c4 --transfer $t > gpurun_out/configs_c4_$t.json 2> gpurun_out/configs_c4_$t.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/configs_c4_$t.json"))
for q in d.get("c4",[]): print("c4 $t", q["batch_rows"], q["query"], round(q["wall_ms"],3), "ms", q["batches_transferred"], "/", q["batches_ideal"], round(q["stream_bytes_total"]/q["wall_ms"]/1e6,1), "GB/s of input")
PY
done
