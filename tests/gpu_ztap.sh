#!/bin/bash
# 2x2 form without loading the zero weight taps: short timeouts first (a wrong expect_tx byte count would hang the kernel)
set -o pipefail
TAG=${1:-ztap}
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stride2" -p no:cacheprovider 2>&1 | tail -3 || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_detector.py -x -q -m gpu -k "every_layer or end_to_end or detector or tilings" -p no:cacheprovider 2>&1 | tail -3
timeout 400 python tests/layer_perf.py 256 3 > gpurun_out/layers_${TAG}.txt 2>&1; head -3 gpurun_out/layers_${TAG}.txt; grep " 3 2 " gpurun_out/layers_${TAG}.txt
PE_TC_ZSKIP=0 timeout 400 python tests/layer_perf.py 256 3 > gpurun_out/layers_${TAG}_nozskip.txt 2>&1; head -1 gpurun_out/layers_${TAG}_nozskip.txt; grep " 3 2 " gpurun_out/layers_${TAG}_nozskip.txt
