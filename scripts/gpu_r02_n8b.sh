set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"])
print("e2e", d["e2e"]["value"]/1e9, d["e2e"].get("rows_per_gpu"), d["e2e"].get("h2d_gbs"))
for q in d["c5"]["queries"]: print(q["label"], q["device_ms"], q["frac"], q["gather"]["ms"], q["gather"]["gbs"], q["parity"][:40])
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"]/1e9)
for q in d["c5"]["queries"]: print(q["label"], q["device_ms"], q["frac"], q["gather"]["ms"], q["gather"]["gbs"])
PY
