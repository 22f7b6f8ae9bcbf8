#!/bin/bash
set -o pipefail
TAG=${1:-ll}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${2:-2700} --launch-count 640 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_$TAG.csv --by-grid > gpurun_out/launches_$TAG.md; head -34 gpurun_out/launches_$TAG.md
