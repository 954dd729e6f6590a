#!/bin/bash
# Multi-GPU session, short: default bench + reference arm + ycb cost-sharded: tools/gpu_multi_short.sh <tag> <N>
TAG=${1:-r04d}; N=${2:-8}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
timeout 900 $RUN --steps 5 --warmup 3 2> gpurun_out/${TAG}_n${N}_bench.err | tee gpurun_out/${TAG}_n${N}_bench.json | cut -c1-200
timeout 600 $RUN --impl reference --steps 2 --warmup 1 2>> gpurun_out/${TAG}_n${N}_bench.err | tee gpurun_out/${TAG}_n${N}_bench_reference.json | cut -c1-200
timeout 900 $RUN --workload ycb --frames 512 --steps 3 --warmup 3 --no-cpu --shard cost 2>> gpurun_out/${TAG}_n${N}_bench.err | tee gpurun_out/${TAG}_n${N}_bench_ycb_cost.json | cut -c1-200
tail -3 gpurun_out/${TAG}_n${N}_bench.err
