#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py > gpurun_out/tc_bringup4.log 2>&1; grep -E "TC  |FAIL|rror|timeout" gpurun_out/tc_bringup4.log | head -30
for dbg in 0 127; do echo "=== PE_TC_DBG=$dbg"; PE_TC_DBG=$dbg timeout 300 python tests/layer_perf.py 64 2 2>&1 | head -16; done > gpurun_out/exp_dbg4.txt 2>&1
cat gpurun_out/exp_dbg4.txt
