#!/bin/bash
# ViTPose-B: parity tests, then per-op times with the tiled and the row-wise attention kernel
set -o pipefail
TAG=${1:-vit}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vitpose.py -x -q -m gpu -p no:cacheprovider 2>&1 | tail -5
timeout 400 python tests/vit_perf.py 256 3 > gpurun_out/vit_${TAG}_tiled.txt 2>&1; cat gpurun_out/vit_${TAG}_tiled.txt
PE_ATT_ROWWISE=1 timeout 400 python tests/vit_perf.py 256 3 > gpurun_out/vit_${TAG}_rowwise.txt 2>&1; head -6 gpurun_out/vit_${TAG}_rowwise.txt
