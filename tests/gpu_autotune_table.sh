#!/bin/bash
# per-layer times plus the measured candidate table of the tiling auto-tuner (PE_TC_VERBOSE=2), 256 crops
set -o pipefail
mkdir -p gpurun_out
TAG=${1:-r02}
PE_TC_VERBOSE=2 timeout 600 python tests/layer_perf.py 256 2 > gpurun_out/layers_$TAG.txt 2> gpurun_out/tune_$TAG.log
head -30 gpurun_out/layers_$TAG.txt
grep "conv_tc tune" gpurun_out/tune_$TAG.log | grep -E "kind=3 Cin=(48|96|192|384) Cout=(48|96|192|384) " | sort -u | cut -c14-150
