#!/bin/bash
TAG=${1:-it}
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py 12 13 14 15 16 3 2>&1 | grep -E "TC  |FAIL|rror|timeout" | awk '{print $1,$2,$(NF-3),$(NF-2),$(NF-1)}'
timeout 300 python tests/layer_perf.py 128 2 > gpurun_out/layers_$TAG.txt 2>&1; head -${LINES_SHOW:-30} gpurun_out/layers_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -E "worst layers|keypoint \|dx\||passed|failed|Error|error|assert" | cut -c1-300 > gpurun_out/pytest_$TAG.log; cat gpurun_out/pytest_$TAG.log
