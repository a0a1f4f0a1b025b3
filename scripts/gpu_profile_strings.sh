python scripts/bench_configs.py --which c3 --c3-rows 50000000 --reps 0 > gpurun_out/c3_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k 'regex:string_gather' -c 3 -o gpurun_out/prof_str -f python scripts/bench_configs.py --which c3 --c3-rows 50000000 --reps 0 > gpurun_out/ncu_str.log 2>&1; echo rc=$?
python scripts/ncu_top.py gpurun_out/prof_str.ncu-rep 14 > gpurun_out/prof_str.txt 2>&1
