#!/bin/bash
# the driver's round-end sequence: pytest -m gpu, smoke(), bench.py
set -o pipefail
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_full_$TAG.log 2>&1; echo "pytest -m gpu rc=$?" | tee gpurun_out/pytest_rc_$TAG.txt; tail -5 gpurun_out/pytest_full_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
