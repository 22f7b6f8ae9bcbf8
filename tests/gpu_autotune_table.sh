#!/bin/bash
set -o pipefail
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py 3 4 6 8 13 2>&1 | grep -E "TC  |FAIL|rror|timeout" | awk '{print $1,$2,$(NF-3),$(NF-2),$(NF-1)}'
PE_TC_VERBOSE=2 timeout 300 python tests/layer_perf.py 128 2 > gpurun_out/layers_tune.txt 2> gpurun_out/tune.log
head -14 gpurun_out/layers_tune.txt
grep "conv_tc tune" gpurun_out/tune.log | grep -E "Cin=(48|96|192|384) Cout=(48|96|192|384) ks=3" | sort -u | cut -c14-130
