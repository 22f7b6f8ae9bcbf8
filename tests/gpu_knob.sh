#!/bin/bash
# usage: gpu_knob.sh VAR v1 v2 ... : per-layer times for each value of an env knob
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  echo "=== $VAR=$v"
  env $VAR=$v timeout 200 python tests/layer_perf.py 128 2 2>&1 | head -${LINES_SHOW:-9}
done | tee gpurun_out/knob_$VAR.txt
