#!/bin/bash
# BASELINE configs[3] sharded over N GPUs (gpurun --gpus N -- 'bash tests/gpu_pipeline_ngpu.sh N TAG [FRAMES] [nobench]'): the wrappers
# under torchrun, NCCL all_gather of detections and keypoints, sharded == unsharded check; then the N-GPU bench line
set -o pipefail
N=${1:-2}; TAG=${2:-r02}; FR=${3:-512}
mkdir -p gpurun_out
timeout 900 python tools/bench_pipeline.py --frames $FR > gpurun_out/pipeline_${TAG}_n1.json 2> gpurun_out/pipeline_${TAG}_n1.err; tail -c 1200 gpurun_out/pipeline_${TAG}_n1.json; tail -3 gpurun_out/pipeline_${TAG}_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/bench_pipeline.py --frames $FR --check \
    > gpurun_out/pipeline_${TAG}_n$N.json 2> gpurun_out/pipeline_${TAG}_n$N.err; tail -c 1500 gpurun_out/pipeline_${TAG}_n$N.json; tail -5 gpurun_out/pipeline_${TAG}_n$N.err
[ "$4" = nobench ] && exit 0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary \
    > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; tail -c 1500 gpurun_out/bench_${TAG}_n$N.json; tail -3 gpurun_out/bench_${TAG}_n$N.err
