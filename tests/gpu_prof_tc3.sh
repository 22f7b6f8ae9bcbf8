#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 100 -c 14 -o gpurun_out/prof_tc3b python tests/layer_perf.py 64 1 > gpurun_out/ncu_tc3.log 2>&1
tail -3 gpurun_out/ncu_tc3.log
