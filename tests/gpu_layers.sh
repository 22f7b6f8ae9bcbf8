#!/bin/bash
mkdir -p gpurun_out
for tc in 0 1; do
PE_TEST_TC=$tc timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "every_layer or end_to_end" 2>&1 | grep -E "worst layers|keypoint \|dx\||passed|failed" | cut -c1-900
done > gpurun_out/layers.log 2>&1
cat gpurun_out/layers.log
