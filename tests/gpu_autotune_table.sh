#!/bin/bash
# per-layer times with the auto-tuned plans, the chosen plan per layer shape, and the measured candidate table (PE_TC_VERBOSE=2), 256 crops
set -o pipefail
mkdir -p gpurun_out
TAG=${1:-r02}
PE_TC_VERBOSE=2 timeout 900 python tests/layer_perf.py 256 2 > gpurun_out/layers_$TAG.txt 2> gpurun_out/tune_$TAG.log
head -40 gpurun_out/layers_$TAG.txt
grep "conv_tc plan" gpurun_out/tune_$TAG.log | awk '{print $3,$4,$5,$6,$8,$9,$10,$11,$12,$13,$14,$15,$16}' | sort | uniq -c | sort -rn | head -40
