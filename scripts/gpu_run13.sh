python -m pytest tests/test_synth_checksums.py -m gpu -x -q > gpurun_out/pytest_synth.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_synth.log
