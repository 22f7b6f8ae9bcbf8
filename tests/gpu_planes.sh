#!/bin/bash
# lane-parallel TMA producer (default) against the single-lane one (PE_TC_PLANES=1): bit-identity first (short timeout), then A/B
set -o pipefail
TAG=${1:-pl}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stride2 or tilings" -p no:cacheprovider 2>&1 | tail -3 || exit 1
for n in 32 1; do
PE_TC_PLANES=$n timeout 300 python tests/layer_perf.py 256 3 > gpurun_out/layers_${TAG}_$n.txt 2>&1; echo "== PE_TC_PLANES=$n"; head -12 gpurun_out/layers_${TAG}_$n.txt; grep " 3 2 \| 1 1 " gpurun_out/layers_${TAG}_$n.txt | head -12
done
for n in 32 1; do
PE_TC_PLANES=$n timeout 300 python tests/vit_perf.py 256 3 2>&1 | head -7
done
