#!/bin/bash
# zero-block skipping of the 2x2 (stride-2) form: parity suite, then the bench with the knob off / on (twice each, alternating)
set -o pipefail
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -5
for v in 0 1 0 1; do
PE_TC_ZSKIP=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('PE_TC_ZSKIP=$v', {k:round(d[k],1) for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), d['clocks'], d['parity']['ok'], d['parity']['max_abs_px_well_conditioned'])"
done
