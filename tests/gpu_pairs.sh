#!/bin/bash
# forms of conv_tc (CG = CTA pairs, SETS = epilogue organisation): single layers vs float64 with each form pinned (spin waits
# with bounded counts: a protocol bug traps instead of hanging), the parity suite, per-layer times per form
set -o pipefail
mkdir -p gpurun_out
for cfg in ${BRINGUP:-"PE_TC_CG=1,PE_TC_SETS=3" "PE_TC_CG=2,PE_TC_SETS=3"}; do
  echo "== bring-up $cfg"
  env ${cfg//,/ } PE_TC_POLL_NS=0 PE_TC_VERBOSE=1 timeout 300 python tests/tc_bringup.py 3 4 5 6 7 11 17 18 2>&1 | grep "TC  \|FAIL\|rror\|timeout" | tail -12
done
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for cfg in ${FORMS:-"PE_TC_CG=1,PE_TC_SETS=1" "PE_TC_CG=1,PE_TC_SETS=3" "PE_TC_CG=2,PE_TC_SETS=3"}; do
  echo "== $cfg"
  env ${cfg//,/ } timeout 300 python tests/layer_perf.py 256 2 2>/dev/null | grep -E "forward|conv +(48 +48|96 +96|192 +192|384 +384|64 +64) 3 1"
done
