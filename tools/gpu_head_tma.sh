#!/bin/bash
# TMA form of the head's copy (RCV_HEAD_TMA=1): parity tests, then bandwidth for a few ring shapes.  Usage: tools/gpu_head_tma.sh <tag>
TAG=${1:-head_tma}
mkdir -p gpurun_out; rm -f gpurun_out/${TAG}.txt
RCV_HEAD_TMA=1 timeout 120 python -m pytest tests -m gpu -q -x -k "head" 2>&1 | tail -15 | tee -a gpurun_out/${TAG}.txt
for cfg in "24 3" "24 2" "34 2" "22 6"; do set -- $cfg
  echo -n "tma $1x$2 " | tee -a gpurun_out/${TAG}.txt
  RCV_HEAD_TMA=1 RCV_HEAD_CFG=$1 RCV_HEAD_CTAS=$2 timeout 60 python tools/head_bw.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}.txt
done
