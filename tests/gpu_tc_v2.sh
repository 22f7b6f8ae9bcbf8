#!/bin/bash
mkdir -p gpurun_out
PE_TC_VERBOSE=1 timeout 300 python tests/tc_bringup.py > gpurun_out/tc_bringup2.log 2>&1; grep -E "TC  |FAIL|Error|error|timeout" gpurun_out/tc_bringup2.log | head -30
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -30 > gpurun_out/pytest_tc2.log
cat gpurun_out/pytest_tc2.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc2.json 2> gpurun_out/bench_tc2.err; cat gpurun_out/bench_tc2.json; tail -5 gpurun_out/bench_tc2.err
