python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ring_depths or mixed_density" 2>&1 | tail -2
for sw in 16 32; do for ss in 2 3; do echo "scan_warps=$sw scan_slots=$ss"; python scripts/profile_one.py --rows 1000000000 --scan-warps $sw --scan-slots $ss --reps 3 --thresholds 998,499 2>&1 | tail -2; done; done
