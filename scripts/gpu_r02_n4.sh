set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; echo "bench n4 rc=$?"; tail -3 gpurun_out/r02_bench_n4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n4.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"])
print("e2e", d["e2e"]["value"]/1e9, d["e2e"].get("rows_per_gpu"), d["e2e"].get("h2d_gbs"))
for q in d["c5"]["queries"]: print(q["label"], q["device_ms"], q["frac"], q["gather"]["ms"], q["gather"]["gbs"], q["parity"][:40])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --impl reference --gpus 4 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
