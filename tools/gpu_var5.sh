#!/bin/bash
# parity subset + variants bench: tools/gpu_var5.sh <tag> <variant names...>
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
for v in "$@"; do
  if [ "$v" = default ]; then unset RCV_LIB_PATH; else export RCV_LIB_PATH=$PWD/build/librcvvote_$v.so; fi
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | sed "s/^{/{\"variant\": \"$v\", /" | tee -a gpurun_out/${TAG}_bench_variants.json | cut -c1-150
done
