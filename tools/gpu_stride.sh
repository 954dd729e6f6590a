#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/${TAG}_pytest.txt
for v in 0 1 0 1; do
  RCV_MIN_STRIDE=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | sed "s/^{/{\"min_stride\": $v, /" | tee -a gpurun_out/${TAG}_bench_variants.json | cut -c1-110
done
