#!/bin/bash
TAG=$1
mkdir -p gpurun_out
for b in 16 32; do
  RCV_C2_FRAMES=$b timeout 900 python tools/config2_bench.py 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_config2_batches.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['frames'], [(v['producer'][:40], v['frames_per_s'], v['stage_ms']) for v in d['variants']])"
done
