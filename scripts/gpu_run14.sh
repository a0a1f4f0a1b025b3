python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu14.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu14.log
python -c "import __graft_entry__ as g; g.smoke()"
