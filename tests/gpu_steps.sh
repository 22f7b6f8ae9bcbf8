#!/bin/bash
mkdir -p gpurun_out
for ms in 9 12; do
echo "=== PE_TC_MAXSTEPS=$ms"
PE_TC_MAXSTEPS=$ms timeout 300 python tests/layer_perf.py 128 2 2>&1 | head -12
PE_TC_MAXSTEPS=$ms timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -s 2>&1 | grep -E "worst layers|keypoint \|dx\||passed|failed|Error|error|assert" | cut -c1-260
done
