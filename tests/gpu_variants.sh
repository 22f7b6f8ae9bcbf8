#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py 1 3 4 5 8 13 2>&1 | grep -E "TC  |FAIL|rror|timeout" | awk '{print $1,$2,$(NF-3),$(NF-2),$(NF-1)}'
for v in "$@"; do
  echo "=== $v"
  env $v timeout 200 python tests/layer_perf.py 128 2 2>&1 | head -${LINES_SHOW:-16}
done | tee gpurun_out/variants.txt
