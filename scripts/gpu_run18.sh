python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu16.log
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv
