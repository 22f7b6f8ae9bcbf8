mkdir -p gpurun_out; python tests/layer_perf.py 64 3 2>&1 | tee gpurun_out/layer_perf.txt
