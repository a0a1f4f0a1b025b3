set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stream or exact or sharded" > gpurun_out/r02_pytest_stream.log 2>&1; echo "pytest-stream rc=$?"; tail -5 gpurun_out/r02_pytest_stream.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_bench_n1.err
head -c 3000 gpurun_out/r02_bench_n1.json
./scripts/microbench_sector 0 > gpurun_out/r02_sector_default.txt 2>&1
./scripts/microbench_sector 32 > gpurun_out/r02_sector_gran32.txt 2>&1
ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_sector_default_ncu.csv ./scripts/microbench_sector 0 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_sector_gran32_ncu.csv ./scripts/microbench_sector 32 > /dev/null 2>&1
tail -3 gpurun_out/r02_sector_default.txt
