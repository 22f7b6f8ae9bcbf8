#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py 8 9 2>&1 | grep -E "TC  |FAIL|rror|timeout" | awk '{print $1,$2,$(NF-3),$(NF-2),$(NF-1)}'
PE_TC_VERBOSE=2 timeout 300 python tests/layer_perf.py 128 2 > gpurun_out/layers_1x1.txt 2> gpurun_out/tune_1x1.log
head -12 gpurun_out/layers_1x1.txt
grep "conv_tc tune" gpurun_out/tune_1x1.log | grep "Cin=64 Cout=256 ks=1 96x72 res=1" | cut -c14-130
