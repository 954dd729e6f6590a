#!/bin/bash
# Multi-GPU session: tools/gpu_multi.sh <tag> <N>
TAG=${1:-r02m}; N=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N"
timeout 900 $RUN --steps 5 --warmup 3 2> gpurun_out/${TAG}_n${N}_bench.err | tee gpurun_out/${TAG}_n${N}_bench.json | cut -c1-200
timeout 600 $RUN --impl reference --steps 2 --warmup 1 2>> gpurun_out/${TAG}_n${N}_bench.err | tee gpurun_out/${TAG}_n${N}_bench_reference.json | cut -c1-200
timeout 900 $RUN --workload ycb --frames 512 --steps 3 --warmup 3 --no-cpu --shard cost 2>> gpurun_out/${TAG}_n${N}_bench.err | tee gpurun_out/${TAG}_n${N}_bench_ycb_cost.json | cut -c1-200
timeout 900 $RUN --workload ycb --frames 512 --steps 3 --warmup 3 --no-cpu --shard count 2>> gpurun_out/${TAG}_n${N}_bench.err | tee gpurun_out/${TAG}_n${N}_bench_ycb_count.json | cut -c1-200
tail -5 gpurun_out/${TAG}_n${N}_bench.err
nvidia-smi topo -m > gpurun_out/${TAG}_n${N}_topo.txt 2>&1; numactl -H >> gpurun_out/${TAG}_n${N}_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/${TAG}_n${N}_topo.txt
