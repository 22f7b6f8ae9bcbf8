#!/bin/bash
mkdir -p gpurun_out
PE_TC_PROF=1 timeout 300 python tests/layer_perf.py 64 1 2>&1 | grep "conv_tc prof" | grep -E "ntaps=9" | awk '{k=$4" "$5" "$6" "$7" "$8; n[k]++; line[k]=$0} END{for(k in n) print n[k], line[k]}' | cut -c1-330 > gpurun_out/tc_prof.txt
cat gpurun_out/tc_prof.txt
