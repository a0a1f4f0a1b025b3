python scripts/scan_sweep.py 2>&1 | tee gpurun_out/r02_scan_sweep.txt
