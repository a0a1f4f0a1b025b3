timeout 600 python -m pytest tests/test_host_golden.py tests/test_gpu_parity.py -m gpu -x -q -k "join" 2>&1 | tail -30
