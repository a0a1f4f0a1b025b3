for h in 0 1 2 3; do echo "RVL_SCAN_L2=$h"; RVL_SCAN_L2=$h python scripts/scan_sweep.py short 2>&1; done | tee gpurun_out/r02_scan_l2hints.txt
