python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu21.log
python -c "import __graft_entry__ as g; g.smoke()"
