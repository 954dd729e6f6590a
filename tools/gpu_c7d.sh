#!/bin/bash
TAG=$1
mkdir -p gpurun_out
for bo in 1 0; do
  echo "== base offset mode $bo"
  RCV_C7_BASE_OFFSET=$bo timeout 600 python -m pytest tests/test_producer.py -m gpu -x -q -k "conv7_head_kernel" 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_pytest_c7.txt
done
timeout 600 python tools/tail_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_tail_bench.json
