ncu --set full --import-source on --clock-control none -k regex:compact_dense -c 2 -o gpurun_out/prof_dense -f python scripts/profile_one.py --rows 1000000000 --plan two_pass --thresholds 899,99 > gpurun_out/ncu_dense.log 2>&1
python scripts/ncu_top.py gpurun_out/prof_dense.ncu-rep 14 > gpurun_out/prof_dense.txt 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
