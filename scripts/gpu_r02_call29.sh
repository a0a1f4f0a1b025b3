set -x
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
timeout 1500 bash scripts/gpu_r02_profile.sh > gpurun_out/r02_profile.log 2>&1; echo "profile rc=$?"; tail -12 gpurun_out/r02_profile.log
du -sh gpurun_out
