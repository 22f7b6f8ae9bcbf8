#!/bin/bash
# which epilogue feature costs the HRNet layers time: rebuild the library with features compiled out, per-layer times each
mkdir -p gpurun_out
for v in "" "-DPE_TC_NO_SILU" "-DPE_TC_NO_SILU -DPE_TC_NO_RESPOST" "-DPE_TC_NO_SILU -DPE_TC_NO_RESPOST -DPE_TC_NO_RANGECHECK"; do
  echo "== variant: [$v]"
  PE_EXTRA_NVCC_FLAGS="$v" python -m posepipeline_b200.csrc.build > /dev/null 2>&1 || echo build failed
  timeout 300 python tests/layer_perf.py 128 2 2>/dev/null | head -12
done
PE_EXTRA_NVCC_FLAGS="" python -m posepipeline_b200.csrc.build > /dev/null 2>&1
