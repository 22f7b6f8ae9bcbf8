#!/bin/bash
# ncu --set full captures of the SIMT kernels rewritten in round 2: tiled attention (ViTPose-B), stem / fuse / head (HRNet)
set -o pipefail
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:attention_tiled_kernel" --launch-skip 12 --launch-count 2 -o gpurun_out/full_${TAG}_att -f \
    python tests/vit_perf.py 256 1 > gpurun_out/ncu_full_${TAG}_att.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:stem_kernel|fuse_kernel|head_kernel" --launch-skip 25 --launch-count 12 -o gpurun_out/full_${TAG}_simt -f \
    python tests/layer_perf.py 256 1 > gpurun_out/ncu_full_${TAG}_simt.log 2>&1
for name in att simt; do
  ncu -i gpurun_out/full_${TAG}_$name.ncu-rep --page raw --csv > gpurun_out/full_${TAG}_$name.csv 2>/dev/null
  python profiles/ncu_summary.py gpurun_out/full_${TAG}_$name.csv > gpurun_out/full_${TAG}_$name.md; head -22 gpurun_out/full_${TAG}_$name.md | cut -c1-400
  rm -f gpurun_out/full_${TAG}_$name.ncu-rep
done
