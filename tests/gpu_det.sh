#!/bin/bash
# detector / tracker parity + HRNet per-layer timing (regression check of the conv_tc host rewrite)
set -o pipefail
TAG=${1:-det}
mkdir -p gpurun_out
timeout 1800 python3 -m pytest tests/test_gpu_detector.py -x -q -s -m gpu -p no:cacheprovider > gpurun_out/pytest_det_$TAG.log 2>&1; echo "pytest detector rc=$?"
grep -E "detector worst|detections img|all-priors|bytetrack on video|passed|failed|Error|error|assert" gpurun_out/pytest_det_$TAG.log | cut -c1-600 | tail -30
PE_TC_VERBOSE=2 timeout 300 python tests/layer_perf.py 128 2 > gpurun_out/layers_$TAG.txt 2> gpurun_out/tune_$TAG.log; head -22 gpurun_out/layers_$TAG.txt
grep "conv_tc tune" gpurun_out/tune_$TAG.log | grep -E "Cin=(48|96|192) Cout=(48|96|192) %?" | grep "kind=3" | sort -u | cut -c14-150 | head -60
