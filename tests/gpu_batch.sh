#!/bin/bash
mkdir -p gpurun_out
for mc in 16 32 64 128 256; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --max-crops $mc 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('max_crops', d['config']['internal_batch_crops'], 'crops/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'convTF', round(d['roofline']['achieved'],1), 'launches', d['gpu_launches'])"
done | tee gpurun_out/batch_sweep.txt
