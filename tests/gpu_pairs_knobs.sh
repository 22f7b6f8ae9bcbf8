#!/bin/bash
# CTA-pair form pinned, wait-policy knobs: per-layer-shape times of the 3x3 layers
mkdir -p gpurun_out
for cfg in "PE_TC_CG=1 PE_TC_POLL_NS=-1000" "PE_TC_CG=2 PE_TC_POLL_NS=-1000" "PE_TC_CG=2 PE_TC_POLL_NS=0" "PE_TC_CG=2 PE_TC_POLL_NS=32" "PE_TC_CG=2 PE_TC_POLL_NS=-200" "PE_TC_CG=1 PE_TC_POLL_NS=0"; do
  echo "== $cfg"
  env $cfg timeout 300 python tests/layer_perf.py 256 2 2>/dev/null | grep -E "forward|conv +(48 +48|96 +96|192 +192|384 +384|64 +64) 3 1"
done
