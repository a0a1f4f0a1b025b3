set -x
# launch list of the bench command (after the same command exited 0 without ncu)
python bench.py --steps 2 --warmup 3 --no-e2e --no-configs --no-golden --cpu-rows 0 --verify-rows 0 > gpurun_out/r02_b_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-configs --no-golden --cpu-rows 0 --verify-rows 0 > gpurun_out/r02_ncu_launches.log 2>&1
# ncu --set full: two-pass kernels at the bench size
python scripts/profile_one.py --rows 1000000000 > gpurun_out/r02_plain1b.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k 'regex:predicate_scan|compact_dense|gather_sparse' -c 12 -o gpurun_out/r02_prof_twopass -f python scripts/profile_one.py --rows 1000000000 > gpurun_out/r02_ncu_full_1b.log 2>&1
python scripts/ncu_top.py gpurun_out/r02_prof_twopass.ncu-rep 8 > gpurun_out/r02_ncu_full_1b_twopass.txt 2>&1
python scripts/make_traffic_json.py gpurun_out/r02_prof_twopass.ncu-rep gpurun_out/r02_traffic.json
ncu -i gpurun_out/r02_prof_twopass.ncu-rep --page raw --csv 2>/dev/null | gzip > gpurun_out/r02_ncu_full_1b_twopass_raw.csv.gz
rm -f gpurun_out/r02_prof_twopass.ncu-rep     # gpurun brings back at most 64 MiB: the summaries travel, the reports do not
# configs[4] shard: chunk kernel (50 %) and the two-pass kernels (10 %)
cat > /tmp/c5_one.py <<'PY'
import sys
sys.path.insert(0, '.')
from rivulus_b200 import capi
ctx = capi.Context(0)
t = ctx.gen_batch([(capi.SYNTH_KEY1000, 0, 0), (capi.SYNTH_F64, 1, 0), (capi.SYNTH_BOOL, 2, 0)], 500_000_000, 3_500_000_000)
ctx.profile_enable(True)
for thr in (899, 499):
    o = ctx.filter_project(t, capi.predicate(0, ">", thr), [0, 1, 2])
    print(thr, o.num_rows(), ctx.profile_read_launches()); o.release()
PY
python /tmp/c5_one.py > gpurun_out/r02_c5_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k 'regex:chunk_filter|compact_bits|predicate_scan|compact_dense|gather_sparse' -c 8 -o gpurun_out/r02_prof_c5 -f python /tmp/c5_one.py > gpurun_out/r02_ncu_c5.log 2>&1
python scripts/ncu_top.py gpurun_out/r02_prof_c5.ncu-rep 8 > gpurun_out/r02_ncu_full_c5.txt 2>&1
rm -f gpurun_out/r02_prof_c5.ncu-rep
# configs[2] batch: string kernels
python scripts/c3_one.py --kernel 3 > gpurun_out/r02_c3_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k 'regex:string_' -c 4 -o gpurun_out/r02_prof_str -f python scripts/c3_one.py --kernel 3 > gpurun_out/r02_ncu_str.log 2>&1
python scripts/ncu_top.py gpurun_out/r02_prof_str.ncu-rep 8 > gpurun_out/r02_ncu_full_c3_strings.txt 2>&1
rm -f gpurun_out/r02_prof_str.ncu-rep
ls -la gpurun_out/
