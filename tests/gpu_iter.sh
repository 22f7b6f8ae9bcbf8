#!/bin/bash
# quick iteration: single-layer bring-up (correctness), per-layer times, then the parity suite
TAG=${1:-it}
mkdir -p gpurun_out
PE_TC_VERBOSE=1 timeout 300 python tests/tc_bringup.py 2>&1 | grep -E "TC  |FAIL|rror|timeout|conv_tc plan" | awk '{print $1,$2,$3,$4,$5,$6,$7,$8,$9,$10,$(NF-3),$(NF-2),$(NF-1)}' > gpurun_out/bringup_$TAG.txt; cat gpurun_out/bringup_$TAG.txt
timeout 300 python tests/layer_perf.py 128 2 > gpurun_out/layers_$TAG.txt 2>&1; head -24 gpurun_out/layers_$TAG.txt
if [ "$2" == "full" ]; then
timeout 1200 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "worst layers|keypoint \|dx\||passed|failed|Error|error|assert" | cut -c1-400 > gpurun_out/pytest_$TAG.log; cat gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['clocks'])"; tail -3 gpurun_out/bench_$TAG.err
fi
