#!/bin/bash
mkdir -p gpurun_out
for ns in 0 -200 -2000 -20000; do
PE_TC_POLL_NS=$ns timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('poll_ns=$ns', {k:round(d[k],1) for k in ('value','ms_per_step')}, round(d['e2e']['value'],1), d['clocks'])"
done
