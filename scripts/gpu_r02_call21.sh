timeout 900 python -m pytest tests/test_host_golden.py -m gpu -q -k "csv" 2>&1 | tail -30
