#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1; do
  echo "=== PE_TC_BO_MODE=$mode" 
  PE_TC_BO_MODE=$mode PE_TC_VERBOSE=1 timeout 300 python tests/tc_bringup.py 2>&1 | tail -45
done > gpurun_out/tc_bringup.log 2>&1
cat gpurun_out/tc_bringup.log
