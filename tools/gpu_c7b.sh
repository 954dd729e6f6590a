#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 600 python tools/tail_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_tail_bench.json
timeout 900 python tools/config2_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_config2.json | cut -c1-1500
RCV_TAIL_IMAGES=6 RCV_TAIL_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv7_head -s 3 -c 1 -f -o gpurun_out/${TAG}_conv7head \
  python tools/tail_bench.py > gpurun_out/${TAG}_ncu_c7.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_c7.log
