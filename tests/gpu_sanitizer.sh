#!/bin/bash
# compute-sanitizer over single conv_tc layers (every kind, every form: one-CTA / CTA-pair x the four epilogue organisations) and
# over smoke(): memcheck (out-of-bounds / misaligned shared + global accesses, invalid barrier use) and synccheck.
# Summary -> gpurun_out/sanitizer_$TAG.txt (copied to profiles/).
TAG=${1:-r02}
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_$TAG.txt
echo "# compute-sanitizer $(compute-sanitizer --version | head -1), $(date -u +%F)" > $OUT
for tool in memcheck synccheck; do
  FORMS="PE_TC_CG=1,PE_TC_SETS=1 PE_TC_CG=2,PE_TC_SETS=1 PE_TC_CG=1,PE_TC_SETS=2 PE_TC_CG=2,PE_TC_SETS=3 PE_TC_CG=2,PE_TC_SETS=4"
  [ $tool = synccheck ] && FORMS="PE_TC_CG=1,PE_TC_SETS=1 PE_TC_CG=2,PE_TC_SETS=4"
  for cfg in $FORMS; do
    echo "== $tool $cfg: tc_bringup cases 3 4 8 9 13 17 (3x3 +res, 1x1 +res, 1x1, stride 2, multi-tile)" >> $OUT
    env ${cfg//,/ } PE_TC_POLL_NS=0 timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tests/tc_bringup.py 3 4 8 9 13 17 2>&1 \
      | grep -E "TC  |ERROR SUMMARY|Invalid|Misaligned|Barrier error|hazard|FAIL" | cut -c1-200 >> $OUT
  done
done
echo "== memcheck smoke() (crop, stem, 291 conv_tc launches, fuse, head, decode; cost-model tilings, no auto-tune runs)" >> $OUT
PE_TC_AUTOTUNE=0 timeout 1500 compute-sanitizer --tool memcheck --print-limit 5 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -E "smoke|ERROR SUMMARY|Invalid|Misaligned|Barrier error" | cut -c1-200 >> $OUT
cat $OUT
