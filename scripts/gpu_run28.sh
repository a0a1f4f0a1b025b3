SEL='strings_long or golden_boolean or stream_zero_copy or test_golden_record_batch_filter or test_limit_semantics or test_all_null or test_sliced_views or test_concat_matches or test_take_matches'
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/san_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/sanitizer_memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/san_run.log 2>&1; echo "sanitizer rc=$?"
tail -3 gpurun_out/san_plain.log; tail -3 gpurun_out/san_run.log; tail -15 gpurun_out/sanitizer_memcheck.log
