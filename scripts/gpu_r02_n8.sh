set -x
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r02_bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["value"]/1e9, d["ms_per_step"], d["roofline"]["frac"])
print("e2e", d["e2e"]["value"]/1e9, d["e2e"].get("rows_per_gpu"), d["e2e"].get("h2d_gbs"))
c5=d["c5"]
for q in c5["queries"]: print(q["label"], q["device_ms"], q["frac"], q["gather"]["ms"], q["gather"]["gbs"], q["parity"][:40])
PY
for n in 8 4 2 1; do timeout 120 python scripts/pcie_scaling.py --gpus $n > gpurun_out/r02_pcie_scaling_n$n.txt 2>&1; tail -2 gpurun_out/r02_pcie_scaling_n$n.txt; done
timeout 120 python scripts/pcie_scaling.py --gpus 8 --affinity > gpurun_out/r02_pcie_scaling_n8_affinity.txt 2>&1; tail -2 gpurun_out/r02_pcie_scaling_n8_affinity.txt
nvidia-smi topo -m > gpurun_out/r02_topo.txt 2>&1; lscpu | head -25 > gpurun_out/r02_lscpu.txt; numactl -H >> gpurun_out/r02_lscpu.txt 2>&1; free -g >> gpurun_out/r02_lscpu.txt
