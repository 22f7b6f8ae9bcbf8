#!/bin/bash
# final visit of a round: the driver's exact pytest command, smoke(), the default bench line, ncu launch list of one step
set -o pipefail
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1500 python3 -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1; echo "pytest -m gpu rc=$?" | tee gpurun_out/pytest_rc_$TAG.txt
tail -3 gpurun_out/pytest_full_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2; echo "smoke rc=$?"
( time timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err ) 2>&1 | grep real; cat gpurun_out/bench_$TAG.json | cut -c1-1500; tail -3 gpurun_out/bench_$TAG.err
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2700 --launch-count 640 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/ncu_bench_$TAG.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_$TAG.csv --by-grid > gpurun_out/launches_$TAG.md; head -24 gpurun_out/launches_$TAG.md
fi
