#!/bin/bash
# grouped work-item schedule of conv_tc (tc_work_item) + tiled attention: bit-identity / parity, then ViTPose-B per-op times
set -o pipefail
TAG=${1:-grp}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vitpose.py tests/test_gpu_detector.py -x -q -m gpu -k "tilings or vitpose or lifter or detector" -p no:cacheprovider 2>&1 | tail -5
timeout 400 python tests/vit_perf.py 256 3 > gpurun_out/vit_${TAG}_qb96.txt 2>&1; head -9 gpurun_out/vit_${TAG}_qb96.txt
PE_ATT_QB=64 timeout 400 python tests/vit_perf.py 256 3 > gpurun_out/vit_${TAG}_qb64.txt 2>&1; head -4 gpurun_out/vit_${TAG}_qb64.txt
timeout 400 python tests/layer_perf.py 256 3 > gpurun_out/layers_${TAG}.txt 2>&1; head -14 gpurun_out/layers_${TAG}.txt
