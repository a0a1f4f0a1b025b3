timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n2.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"]/1e9)
for q in d["c5"]["queries"]: print(q["label"], q["device_ms"], q["frac"], q["gather"]["ms"], q["gather"]["gbs"], q["parity"][:50])
PY
