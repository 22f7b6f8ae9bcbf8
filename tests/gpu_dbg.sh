#!/bin/bash
mkdir -p gpurun_out
for cfg in "PE_TC_CG=1 PE_TC_SETS=2 PE_TC_AUTOTUNE=0" "PE_TC_CG=1 PE_TC_SETS=2 PE_TC_AUTOTUNE=1"; do
echo "== $cfg"
env $cfg PE_TC_VERBOSE=2 timeout 120 python tests/layer_perf.py 16 1 2>&1 | grep -E "conv_tc|forward|rror" | tail -8 | cut -c1-260
done
