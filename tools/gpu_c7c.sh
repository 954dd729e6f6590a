#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_producer.py -m gpu -x -q -k "conv7 or fused_tail" 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_c7.txt
timeout 600 python tools/tail_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_tail_bench.json
RCV_TAIL_IMAGES=6 RCV_TAIL_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv7_head -s 3 -c 1 -f -o gpurun_out/${TAG}_conv7head \
  python tools/tail_bench.py > gpurun_out/${TAG}_ncu_c7.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_c7.log
