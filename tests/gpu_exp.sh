#!/bin/bash
mkdir -p gpurun_out
for dbg in 1 2 12 16 18 3; do echo "=== PE_TC_DBG=$dbg"; PE_TC_DBG=$dbg timeout 300 python tests/layer_perf.py 64 2 2>&1 | grep -E "forward|  48   48 3 1|  96   96 3 1|192  192 3 1|384  384 3 1"; done > gpurun_out/exp_dbg5.txt 2>&1
cat gpurun_out/exp_dbg5.txt
