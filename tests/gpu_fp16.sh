#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py 2>&1 | grep -E "max rel err|FAIL|rror|timeout" | awk '{print $1,$2,$(NF-5),$(NF-4),$(NF-3),$(NF-2),$(NF-1)}' > gpurun_out/fp16_bringup.txt; cat gpurun_out/fp16_bringup.txt | head -40
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "worst layers|keypoint \|dx\||passed|failed|Error|error|assert" | cut -c1-600 | tee gpurun_out/fp16_pytest.txt
timeout 300 python tests/layer_perf.py 64 2 2>&1 | head -16 | tee gpurun_out/fp16_layers.txt
