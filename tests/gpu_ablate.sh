#!/bin/bash
# timing-only ablations of conv_tc (PE_TC_DBG bits: 1 no MMAs, 2 no epilogue global traffic, 4 no weight loads, 8 no activation loads, 16 no TMEM drains)
mkdir -p gpurun_out
for dbg in 0 1 2 16 18 19 12; do
  echo "=== PE_TC_DBG=$dbg"
  PE_TC_DBG=$dbg timeout 200 python tests/layer_perf.py 128 2 2>&1 | head -12
done > gpurun_out/ablate.txt 2>&1
PE_TC_PROF=1 timeout 300 python tests/layer_perf.py 128 1 2>&1 | grep "conv_tc prof" | sort | uniq -c | sort -rn | awk '{ $1=""; print }' | sort -u -t'|' -k1,1 | head -60 > gpurun_out/prof_cycles.txt
cat gpurun_out/ablate.txt; head -40 gpurun_out/prof_cycles.txt
