#!/bin/bash
# Bench every build/librcvvote_*.so variant (device-resident only) + the default library.
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | sed "s/^{/{\"variant\": \"default\", /" | tee gpurun_out/${TAG}_bench_variants.json | cut -c1-140
for so in build/librcvvote_*.so; do
  v=$(basename $so .so); v=${v#librcvvote_}
  RCV_LIB_PATH=$PWD/$so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | sed "s/^{/{\"variant\": \"$v\", /" | tee -a gpurun_out/${TAG}_bench_variants.json | cut -c1-140
done
