#!/bin/bash
TAG=$1
mkdir -p gpurun_out
for v in q40 default q40 default; do
  if [ "$v" = default ]; then unset RCV_LIB_PATH; else export RCV_LIB_PATH=$PWD/build/librcvvote_$v.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"$v\", /" | tee -a gpurun_out/${TAG}_bench_variants.json | cut -c1-120
done
