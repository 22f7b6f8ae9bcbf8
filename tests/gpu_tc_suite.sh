#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/pytest_tc.log
cat gpurun_out/pytest_tc.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; cat gpurun_out/bench_tc.json; tail -5 gpurun_out/bench_tc.err
