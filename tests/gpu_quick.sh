#!/bin/bash
set -o pipefail
TAG=${1:-it}
mkdir -p gpurun_out
timeout 300 python tests/tc_bringup.py 13 3 2>&1 | grep -E "TC  |FAIL|rror|timeout" | awk '{print $1,$2,$(NF-3),$(NF-2),$(NF-1)}'
timeout 300 python tests/layer_perf.py 128 2 > gpurun_out/layers_$TAG.txt 2>&1; head -${LINES_SHOW:-30} gpurun_out/layers_$TAG.txt
timeout 1200 python -m pytest tests -m gpu -q -s -x 2>&1 | grep -E "worst layers|keypoint \|dx\||passed|failed|Error|error|assert" | cut -c1-300 > gpurun_out/pytest_$TAG.log; cat gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['profiled_pass_ms_per_step'], d['clocks'])"; tail -3 gpurun_out/bench_$TAG.err
PE_GRAPH=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('PE_GRAPH=0', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['clocks'])"
