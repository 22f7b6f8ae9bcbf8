#!/bin/bash
# only the ncu --set full captures of tests/gpu_profile.sh (steady-state launches: the skips jump over the auto-tuner's candidate runs)
set -o pipefail
TAG=${1:-r02}
mkdir -p gpurun_out
I='\\(int\\)'
for spec in ${SPECS:-"6, 2, 9, 1, [12], [1234]:750:4:k48"} "6, 1, 9, 1, 2, [1234]:700:6:k96"; do
  IFS=: read tpl skip cnt name <<< "$spec"
  tpl=$(echo "$tpl" | sed -E "s/(\\[[0-9]+\\]|[0-9]+)/$I\\1/g")
  timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:conv_tc_kernel<$tpl>" --launch-skip $skip --launch-count $cnt -o gpurun_out/full_${TAG}_$name -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/ncu_full_${TAG}_$name.log 2>&1
  ncu -i gpurun_out/full_${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_$name.csv 2>/dev/null
  python profiles/ncu_summary.py gpurun_out/full_${TAG}_$name.csv $( [ $name = k48 ] && echo --json gpurun_out/top_kernel_traffic_$TAG.json ) > gpurun_out/full_${TAG}_$name.md; head -16 gpurun_out/full_${TAG}_$name.md | cut -c1-330
  rm -f gpurun_out/full_${TAG}_$name.ncu-rep
done
