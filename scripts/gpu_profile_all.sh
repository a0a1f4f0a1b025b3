set -x
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-e2e --cpu-rows 0 --verify-rows 0 > gpurun_out/b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-e2e --cpu-rows 0 --verify-rows 0 > gpurun_out/ncu_launches.log 2>&1
python scripts/profile_one.py --rows 1000000000 > gpurun_out/plain1b.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k 'regex:predicate_scan|compact_dense|gather_sparse' -c 12 -o gpurun_out/prof_r01_twopass -f python scripts/profile_one.py --rows 1000000000 > gpurun_out/ncu_full_1b.log 2>&1
python scripts/ncu_top.py gpurun_out/prof_r01_twopass.ncu-rep 10 > gpurun_out/prof_r01_twopass.txt 2>&1
python scripts/bench_configs.py --which c5 --c5-rows 500000000 --reps 0 > gpurun_out/c5_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k 'regex:compact_bits|compact_dense' -c 4 -o gpurun_out/prof_r01_c5 -f python scripts/bench_configs.py --which c5 --c5-rows 500000000 --reps 0 > gpurun_out/ncu_c5_full.log 2>&1
python scripts/ncu_top.py gpurun_out/prof_r01_c5.ncu-rep 10 > gpurun_out/prof_r01_c5.txt 2>&1
python scripts/bench_configs.py --which c3 --c3-rows 50000000 --reps 0 > gpurun_out/c3_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k 'regex:string_gather' -c 3 -o gpurun_out/prof_r01_str -f python scripts/bench_configs.py --which c3 --c3-rows 50000000 --reps 0 > gpurun_out/ncu_str.log 2>&1
python scripts/ncu_top.py gpurun_out/prof_r01_str.ncu-rep 12 > gpurun_out/prof_r01_str.txt 2>&1
head -c 1200 gpurun_out/bench_full.json
