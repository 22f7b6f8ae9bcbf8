#!/bin/bash
# development visit: selected parity tests (verbose numbers), bench line, lifter bench
set -o pipefail
TAG=${1:-step}; shift
mkdir -p gpurun_out
timeout 1500 python3 -m pytest tests/test_gpu_parity.py -x -q -s -m gpu -p no:cacheprovider "$@" > gpurun_out/pytest_full_$TAG.log 2>&1; echo "pytest rc=$?"
grep -E "UNCONDITIONAL|worst layers|lifter N|passed|failed|Error|error|assert|conv_tc" gpurun_out/pytest_full_$TAG.log | cut -c1-400 | tail -40
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$TAG.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['parity'], d['clocks'])"; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python tools/bench_lifter.py 16384 5 2>&1 | tail -2
PE_LIFTER_TC=0 timeout 300 python tools/bench_lifter.py 16384 5 2>&1 | tail -1
