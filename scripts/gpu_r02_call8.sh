timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunk_plan" > gpurun_out/r02_pytest_chunk3.log 2>&1; echo "pytest-chunk rc=$?"; tail -5 gpurun_out/r02_pytest_chunk3.log
bash scripts/gpu_r02_call7.sh
