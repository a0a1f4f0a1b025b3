for d in 0 1 2 3; do echo "RVL_SCAN_DEBUG=$d"; RVL_SCAN_DEBUG=$d python scripts/scan_sweep.py short 2>&1; done | tee gpurun_out/r02_scan_ablation.txt
