#!/bin/bash
mkdir -p gpurun_out
for mc in 16 32 64 128; do
  echo "=== max_crops=$mc"
  timeout 200 python tests/layer_perf.py $mc 3 2>&1 | head -10
done | tee gpurun_out/batch.txt
