#!/bin/bash
# stride-2 form choice (TMA gather vs s2d copy + 2x2 layer), fuse / head kernels: parity first, then per-layer times A/B
set -o pipefail
TAG=${1:-s2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stride2 or every_layer or end_to_end or halpe or internal_batching" -p no:cacheprovider 2>&1 | tail -5
PE_TC_VERBOSE=1 timeout 400 python tests/layer_perf.py 256 3 > gpurun_out/layers_${TAG}_auto.txt 2> gpurun_out/layers_${TAG}_auto.err
grep "^stride-2" gpurun_out/layers_${TAG}_auto.err
head -45 gpurun_out/layers_${TAG}_auto.txt
PE_TC_S2D=0 timeout 400 python tests/layer_perf.py 256 3 > gpurun_out/layers_${TAG}_gather.txt 2>&1
head -3 gpurun_out/layers_${TAG}_gather.txt; grep " 3 2 \|fuse\|head" gpurun_out/layers_${TAG}_gather.txt
