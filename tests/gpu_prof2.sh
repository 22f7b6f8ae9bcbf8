#!/bin/bash
mkdir -p gpurun_out
for dbg in 31 30; do
echo "=== dbg $dbg"
PE_TC_DBG=$dbg PE_TC_PROF=1 timeout 300 python tests/layer_perf.py 128 1 2>&1 | grep "conv_tc prof" | sed 's/per-CTA cycles //' > gpurun_out/prof_dbg$dbg.txt
for pat in "NC=48 MT=2 TPS=3 nchunk=3 " "NC=96 MT=1 TPS=3 nchunk=6 "; do
  grep "$pat" gpurun_out/prof_dbg$dbg.txt | grep " X " | sed -n 3p; grep "$pat" gpurun_out/prof_dbg$dbg.txt | grep " Y " | sed -n 3p
done | cut -c1-330
done
