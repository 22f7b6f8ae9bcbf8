#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 640 --csv --log-file gpurun_out/launches_tc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --max-crops 64 > gpurun_out/ncu_tc.log 2>&1
tail -2 gpurun_out/ncu_tc.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 60 -c 6 -o gpurun_out/prof_tc python bench.py --steps 1 --warmup 3 --no-cpu-baseline --max-crops 64 > gpurun_out/ncu_tc_full.log 2>&1
tail -2 gpurun_out/ncu_tc_full.log; ls -la gpurun_out
