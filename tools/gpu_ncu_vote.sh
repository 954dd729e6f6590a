#!/bin/bash
# ncu full capture of the vote kernel only.  Usage: tools/gpu_ncu_vote.sh <tag>
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vote -s 3 -c 1 -f -o gpurun_out/${TAG}_vote \
  python bench.py --steps 1 --warmup 1 --frames 256 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_vote.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_vote.log
