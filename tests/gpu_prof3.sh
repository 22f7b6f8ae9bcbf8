#!/bin/bash
mkdir -p gpurun_out
PE_TC_PROF=1 timeout 300 python tests/layer_perf.py 128 1 2>&1 | grep "conv_tc prof" | sed 's/per-CTA cycles //' > gpurun_out/prof7.txt
for pat in "NC=48 MT=2 TAPS=9 KC=1 nchunk=3 " "NC=48 MT=1 TAPS=9 KC=1 nchunk=6 " "NC=48 MT=1 TAPS=9 KC=1 nchunk=12 " "NC=48 MT=1 TAPS=9 KC=1 nchunk=24 " "TAPS=1 KC=4 nchunk=4 "; do
  for r in 0 1; do grep "$pat" gpurun_out/prof7.txt | grep "res=$r" | grep " X " | sed -n 3p; grep "$pat" gpurun_out/prof7.txt | grep "res=$r" | grep " Y " | sed -n 3p; done
done | cut -c1-330
