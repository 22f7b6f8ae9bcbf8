#!/bin/bash
mkdir -p gpurun_out
PE_TC_CG=1 PE_TC_SETS=3 PE_TC_POLL_NS=0 PE_TC_VERBOSE=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tests/tc_bringup.py 17 2>&1 | grep -v "^conv_tc tune" | head -60
