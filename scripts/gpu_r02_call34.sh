timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 0 --print-limit 20 python -m pytest tests/test_host_golden.py -m gpu -x -q -k "known_answers or random_joins or random_csv" > gpurun_out/r02_memcheck_host.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|error" gpurun_out/r02_memcheck_host.log | head -20
tail -5 gpurun_out/r02_memcheck_host.log
