nvidia-smi topo -m 2>/dev/null | head -20 > gpurun_out/topo.txt
for i in 0 1 2 3; do CUDA_VISIBLE_DEVICES=$i scripts/microbench_pcie > gpurun_out/pcie_conc_$i.txt 2>&1 & done
wait
for i in 0 1 2 3; do echo "== GPU $i (4 concurrent)"; grep -E "copy engine\), 2 GiB|148 CTAs x 16 warps, ring depth 3|duplex" gpurun_out/pcie_conc_$i.txt; done
cat gpurun_out/topo.txt
