#!/bin/bash
# the measurement record of a round: bench line, ncu launch list of the steady state, ncu --set full captures of the kernels
# VERDICT r1 names (<6,2,9,1> = 48->48 3x3, <8,1,1,1> = layer1 1x1 + residual) plus the 96/192-channel layers
set -o pipefail
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_all_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/ncu_bench_$TAG.log 2>&1
python - <<PY
import csv
rows = [l for l in open("gpurun_out/launches_all_$TAG.csv") if l.startswith('"')]
hdr, body = rows[0], rows[1:]
# the last forward of the run (the extra e2e step): crop kernel .. decode kernel
idx = [i for i, l in enumerate(body) if "warp_crop_kernel" in l]
last = body[idx[-1]:]
open("gpurun_out/launches_$TAG.csv", "w").write(hdr + "".join(last))
print("launch list: %d launches in total, last step has %d" % (len(body), len(last)))
PY
python profiles/summarize_launches.py gpurun_out/launches_$TAG.csv --by-grid > gpurun_out/launches_$TAG.md; head -24 gpurun_out/launches_$TAG.md
# ncu --set full captures: the 48 -> 48 3x3 layers (one-CTA form), the 96 / 192-channel 3x3 layers (CTA-pair forms, whichever epilogue
# organisation the tuner picked), layer1's 1x1 + residual.  Template arguments: <NG, MT, TAPS, KC, CG, SETS>
I='\\(int\\)'
for spec in "6, 2, 9, 1, [12], [1234]:750:4:k48" "6, 1, 9, 1, 2, [1234]:700:6:k96"; do
  IFS=: read tpl skip cnt name <<< "$spec"
  tpl=$(echo "$tpl" | sed -E "s/(\\[[0-9]+\\]|[0-9]+)/$I\\1/g")          # demangled names read conv_tc_kernel<(int)6, (int)2, ...>
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel<$tpl>" --launch-skip $skip --launch-count $cnt -o gpurun_out/full_${TAG}_$name -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/ncu_full_${TAG}_$name.log 2>&1
  ncu -i gpurun_out/full_${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_$name.csv 2>/dev/null
  python profiles/ncu_summary.py gpurun_out/full_${TAG}_$name.csv $( [ $name = k48 ] && echo --json gpurun_out/top_kernel_traffic_$TAG.json ) > gpurun_out/full_${TAG}_$name.md; head -14 gpurun_out/full_${TAG}_$name.md | cut -c1-260
  rm -f gpurun_out/full_${TAG}_$name.ncu-rep
done
# BASELINE configs[3] through the wrappers on one GPU (N>1: tests/gpu_pipeline_ngpu.sh under gpurun --gpus N)
timeout 900 python tools/bench_pipeline.py --frames 256 > gpurun_out/pipeline_${TAG}_n1.json 2> gpurun_out/pipeline_${TAG}_n1.err; tail -c 1500 gpurun_out/pipeline_${TAG}_n1.json; tail -3 gpurun_out/pipeline_${TAG}_n1.err
