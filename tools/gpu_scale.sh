#!/bin/bash
# default bench at N GPUs only: tools/gpu_scale.sh <tag> <N>
TAG=$1; N=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N"
timeout 900 $RUN --steps 5 --warmup 3 2> gpurun_out/${TAG}_n${N}_bench.err | tee gpurun_out/${TAG}_n${N}_bench.json | cut -c1-200
