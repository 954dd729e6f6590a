#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 600 python tools/dropin_latency.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_dropin_latency.json
