python -m pytest tests/test_host_golden.py -m gpu -x -q > gpurun_out/pytest_host2.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_host2.log
