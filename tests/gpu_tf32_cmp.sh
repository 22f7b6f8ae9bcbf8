#!/bin/bash
set -o pipefail
# the wide-range tf32x3 build of the same sources through the parity suite
PE_PRECISION=tf32 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5
