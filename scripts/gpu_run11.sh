python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu13.log
