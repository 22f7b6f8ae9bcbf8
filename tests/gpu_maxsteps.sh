#!/bin/bash
# accumulation-group length (PE_TC_MAXSTEPS) against keypoint error and speed
set -o pipefail
mkdir -p gpurun_out
for ms in 2 3 6; do
  echo "== PE_TC_MAXSTEPS=$ms"
  PE_TC_MAXSTEPS=$ms PE_SUBRUN=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -p no:cacheprovider -k "end_to_end or halpe or every_layer" 2>&1 | grep -E "UNCONDITIONAL|worst layers|passed|failed" | cut -c1-330
  PE_TC_MAXSTEPS=$ms timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('MAXSTEPS=$ms', {k:round(d[k],1) for k in ('value','ms_per_step')}, round(d['e2e']['value'],1), d['parity'], d['clocks'])"
done
