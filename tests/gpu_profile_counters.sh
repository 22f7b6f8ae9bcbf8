#!/bin/bash
# needs a build with PE_EXTRA_NVCC_FLAGS=-DPE_TC_PROFILE=1: cycle counters of the two MMA warps and the epilogue sections per
# layer shape, for the forms named in FORMS ("cg,sets" pairs)
set -o pipefail
mkdir -p gpurun_out
for f in ${FORMS:-1,1 2,1 2,2}; do
cg=${f%,*}; sets=${f#*,}
PE_TC_CG=$cg PE_TC_SETS=$sets PE_TC_PROF=1 timeout 300 python tests/layer_perf.py 128 1 2>&1 | grep "conv_tc prof" | sed 's/per-CTA cycles //' > gpurun_out/prof_cg${cg}_s$sets.txt
echo "=== CG=$cg SETS=$sets"
for pat in "NC=48 MT=2 TAPS=9 KC=1 nchunk=3 " "NC=96 MT=1 TAPS=9 KC=1 nchunk=6 " "NC=96 MT=1 TAPS=9 KC=1 nchunk=12 "; do
  for r in 0 1; do for role in X Y E; do grep "$pat" gpurun_out/prof_cg${cg}_s$sets.txt | grep "res=$r" | grep "prof $role " | tail -1; done; done
done | cut -c14-420
done
