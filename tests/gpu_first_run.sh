#!/bin/bash
# first GPU bring-up: SIMT-only parity then the default (tensor-core when available) path
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
PE_TEST_TC=0 timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_simt.log
cat gpurun_out/pytest_simt.log
