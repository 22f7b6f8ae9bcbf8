#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/tmem_bench > gpurun_out/tmem_bench.txt 2>&1; cat gpurun_out/tmem_bench.txt
PE_TC_PROF=1 timeout 300 python tests/layer_perf.py 128 1 2>&1 | grep "conv_tc prof" | sed 's/per-CTA cycles //' | awk '{k=$3" "$4" "$5" "$6" "$7" "$8; sub(/.*res=/,"res=",k2); if(!(k in s)){s[k]=0} s[k]++; if(s[k]==2) print}' > gpurun_out/prof_cycles.txt
cat gpurun_out/prof_cycles.txt
