#!/bin/bash
mkdir -p gpurun_out
PE_TC_PROF=1 timeout 300 python tests/layer_perf.py 128 1 2>&1 | grep "conv_tc prof" | sed 's/per-CTA cycles //' > gpurun_out/prof_all.txt
for pat in "NC=48 MT=2 TPS=3 nchunk=3 " "NC=96 MT=1 TPS=3 nchunk=6 " "NC=96 MT=1 TPS=3 nchunk=12 " "NC=96 MT=1 TPS=3 nchunk=24 " "NC=128 MT=1 TPS=1 nchunk=4 " "NC=96 MT=1 TPS=2 nchunk=12 ntaps=4 work=3800"; do
  grep "$pat" gpurun_out/prof_all.txt | grep " X " | sed -n 3p; grep "$pat" gpurun_out/prof_all.txt | grep " Y " | sed -n 3p
  grep "$pat" gpurun_out/prof_all.txt | grep "res=1" | grep " X " | sed -n 3p; grep "$pat" gpurun_out/prof_all.txt | grep "res=1" | grep " Y " | sed -n 3p
done | cut -c1-330
for dbg in 1 2 16 19; do
  echo "=== PE_TC_DBG=$dbg"
  PE_TC_DBG=$dbg timeout 200 python tests/layer_perf.py 128 2 2>&1 | head -8
done
