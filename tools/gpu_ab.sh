#!/bin/bash
# same-session A/B: tools/gpu_ab.sh <tag> <variant> (variant .so in build/ against the default build; env RCV_FULL_WINDOW passed through)
TAG=$1; V=$2
mkdir -p gpurun_out
for v in $V default $V default; do
  if [ "$v" = default ]; then unset RCV_LIB_PATH; else export RCV_LIB_PATH=$PWD/build/librcvvote_$v.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"$v\", /" | tee -a gpurun_out/${TAG}_bench_variants.json | cut -c1-120
done
unset RCV_LIB_PATH
RCV_FULL_WINDOW=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"default_full_window\", /" | tee -a gpurun_out/${TAG}_bench_variants.json | cut -c1-120
