#!/bin/bash
mkdir -p gpurun_out
for r in 1 2 3; do
PE_TC_VERBOSE=2 timeout 300 python tests/layer_perf.py 128 2 > gpurun_out/layers_tune$r.txt 2> gpurun_out/tune$r.log
head -12 gpurun_out/layers_tune$r.txt
grep "conv_tc plan" gpurun_out/tune$r.log | grep -E "Cin=(48|96|192|384) Cout=(48|96|192|384) ks=3" | sort -u | cut -c1-95
done
grep "conv_tc tune" gpurun_out/tune3.log | grep -E "Cin=(96|192) Cout=(96|192) ks=3" | cut -c14-130
