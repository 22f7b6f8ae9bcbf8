#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py 2>&1 | grep -E "TC  |FAIL" | awk '{print $1,$2,$(NF-3),$(NF-2),$(NF-1)}' | head -20
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "worst layers|keypoint \|dx\||passed|failed|AssertionError|Error" | cut -c1-420
python tests/layer_perf.py 64 2 2>/dev/null| head -12
PE_PRECISION=tf32 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "halpe or end_to_end or every_layer" 2>&1 | grep -E "worst layers|keypoint \|dx\||passed|failed|AssertionError" | cut -c1-300
