#!/bin/bash
# quick visit: conv parity suite, then the bench N times (fresh process each: auto-tune stability) with the chosen 3x3 / 2x2 plans
set -o pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for i in ${RUNS:-1 2 3}; do
PE_TC_VERBOSE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>gpurun_out/check_plan_$i.log | python -c "
import json,sys; d=json.load(sys.stdin); print('run $i', {k:round(d[k],1) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), d['clocks'], d['parity']['ok'], d['parity']['max_abs_px_well_conditioned'])"
grep "conv_tc plan: kind=3" gpurun_out/check_plan_$i.log | awk '{print $4,$5,$6,$8,$9,$10,$11,$13,$14,$15,$16}' | sort | uniq -c | sort -rn | head -9
done
grep "conv_tc plan: kind=2\|drain groups" gpurun_out/check_plan_1.log | awk '{print $3,$4,$5,$6,$8,$9,$10,$11,$12,$13,$14,$15,$16,$17,$18}' | sort | uniq -c | sort -rn | head -24
