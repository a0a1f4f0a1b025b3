timeout 300 python -m pytest tests/test_host_golden.py -m gpu -x -q -k "rb_memory or rb_builder or demo" 2>&1 | tail -6
