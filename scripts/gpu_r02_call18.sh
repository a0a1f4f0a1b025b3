set -x
python bench.py --trace --no-e2e --no-configs --no-golden --cpu-rows 0 2> gpurun_out/r02_trace_a.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('A verify on', d['value']/1e9, d['sweep'][0]['kernel_ms_per_step'])"
grep trace gpurun_out/r02_trace_a.err
python bench.py --trace --no-e2e --no-configs --no-golden --cpu-rows 0 --verify-rows 0 2> gpurun_out/r02_trace_b.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B verify off', d['value']/1e9, d['sweep'][0]['kernel_ms_per_step'])"
grep trace gpurun_out/r02_trace_b.err
python bench.py --trace --no-e2e --no-configs --no-golden --cpu-rows 0 2> gpurun_out/r02_trace_c.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C verify on again', d['value']/1e9, d['sweep'][0]['kernel_ms_per_step'])"
grep trace gpurun_out/r02_trace_c.err
