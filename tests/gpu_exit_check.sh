#!/bin/bash
# GPUTEST gate: the driver's exact pytest command, its exit code (not a grep of its output), then smoke()
set -o pipefail
TAG=${1:-exit}
mkdir -p gpurun_out
timeout 1500 python3 -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1; rc=$?
echo "pytest -m gpu rc=$rc" | tee gpurun_out/pytest_rc_$TAG.txt
tail -15 gpurun_out/pytest_full_$TAG.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; echo "smoke rc=$?"
exit $rc
