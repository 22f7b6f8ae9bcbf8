#!/bin/bash
mkdir -p gpurun_out
PE_TC_VERBOSE=2 timeout 300 python tests/layer_perf.py 128 2 > gpurun_out/layers_tune.txt 2> gpurun_out/tune.log
grep "conv_tc tune" gpurun_out/tune.log | sort -u | sort -t= -k2,2n | cut -c14-140 > gpurun_out/tune_table.txt
head -30 gpurun_out/layers_tune.txt
grep -c "conv_tc tune" gpurun_out/tune.log
