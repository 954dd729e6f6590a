#!/bin/bash
# variants bench: tools/gpu_var4.sh <tag> <variant names...>
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then L=""; else L=$PWD/build/librcvvote_$v.so; fi
  RCV_LIB_PATH=$L timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"$v\", /" | tee -a gpurun_out/${TAG}_bench_variants.json | cut -c1-150
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
