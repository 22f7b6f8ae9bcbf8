#!/bin/bash
set -o pipefail
mkdir -p gpurun_out

PE_TC_PROF=1 timeout 300 python tests/layer_perf.py 128 1 2>&1 | grep "conv_tc prof" | sed 's/per-CTA cycles //' > gpurun_out/prof9.txt
for pat in "NC=48 MT=2 TAPS=9 KC=1 nchunk=3 " "NC=96 MT=1 TAPS=9 KC=1 nchunk=6 " "NC=48 MT=2 TAPS=9 KC=1 nchunk=6 " "TAPS=1 KC=2 nchunk=4 "; do
  for r in 0 1; do for role in X E; do grep "$pat" gpurun_out/prof9.txt | grep "res=$r" | grep "prof $role " | tail -1; done; done
done | cut -c14-330
