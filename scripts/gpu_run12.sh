python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
tail -3 gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_n2.json","gpurun_out/bench_ref_n2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), d["value"], d.get("ms_per_step"), d.get("roofline",{}).get("frac"), d.get("e2e",{}).get("value"))
    except Exception as e: print(f, "ERR", e, open(f).read()[:500])
PY
