#!/bin/bash
set -o pipefail
# one GPU-box visit: parity suite, bench line, per-layer times, ncu launch list of one timed step, one --set full capture
TAG=${1:-cur}
mkdir -p gpurun_out
# the driver's exact command first (exit code of the interpreter counts: round 1 segfaulted AFTER "15 passed")
timeout 1500 python3 -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_full_$TAG.log 2>&1; echo "pytest -m gpu rc=$?" | tee gpurun_out/pytest_rc_$TAG.txt
grep -E "worst layers|keypoint \|dx\||passed|failed|Error|error|assert" gpurun_out/pytest_full_$TAG.log | cut -c1-500 > gpurun_out/pytest_$TAG.log
cat gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 300 python tests/layer_perf.py 128 2 > gpurun_out/layers_$TAG.txt 2>&1; head -40 gpurun_out/layers_$TAG.txt
if [ "$2" != "noncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2700 --launch-count 640 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_$TAG.csv --by-grid > gpurun_out/launches_$TAG.md; head -30 gpurun_out/launches_$TAG.md
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc --launch-skip 2300 --launch-count 14 -o gpurun_out/full_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; tail -3 gpurun_out/ncu_full_$TAG.log
fi
