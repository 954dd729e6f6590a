#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_producer.py -m gpu -x -q -k "conv7 or fused_tail" 2>&1 | tail -25 | tee gpurun_out/${TAG}_pytest_c7.txt
