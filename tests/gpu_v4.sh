#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py > gpurun_out/tc_bringup4.log 2>&1; grep -E "TC  |FAIL|rror|timeout" gpurun_out/tc_bringup4.log | awk '{print $1,$2,$12,$13,$14,$15,$16}' | head -30
for dbg in 0; do echo "=== PE_TC_DBG=$dbg"; PE_TC_VERBOSE=1 PE_TC_DBG=$dbg timeout 300 python tests/layer_perf.py 64 2 2>&1 | grep -E "forward|kind|^conv|^fuse|^stem|^head|ks=3 " | sort | uniq | head -50; done > gpurun_out/exp_dbg4.txt 2>&1
cat gpurun_out/exp_dbg4.txt
