#!/bin/bash
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --no-tc --steps 3 --warmup 3 > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; cat gpurun_out/bench_simt.json; tail -3 gpurun_out/bench_simt.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2500 -c 700 --csv --log-file gpurun_out/launches_simt.csv python bench.py --no-tc --steps 1 --warmup 3 --no-cpu-baseline --max-crops 32 > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log
