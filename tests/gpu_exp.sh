#!/bin/bash
mkdir -p gpurun_out
for dbg in 30 14; do echo "=== PE_TC_DBG=$dbg"; PE_TC_DBG=$dbg PE_TC_VERBOSE=1 timeout 300 python tests/layer_perf.py 64 2 2>&1 | grep -E "forward|  48   48 3 1|  96   96 3 1|192  192 3 1|384  384 3 1|plan: Cin=48 Cout=48 ks=3|plan: Cin=96 Cout=96 ks=3|plan: Cin=192 Cout=192 ks=3|plan: Cin=384 Cout=384 ks=3" | sort | uniq; done > gpurun_out/exp_dbg7.txt 2>&1
cat gpurun_out/exp_dbg7.txt
