#!/bin/bash
set -o pipefail
VAR=$1; shift
for v in "$@"; do
env $VAR=$v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('$VAR=$v', {k:round(d[k],1) for k in ('value','ms_per_step')}, round(d['e2e']['value'],1), d['clocks'])"
done
