#!/bin/bash
# e2e host-path check: tools/gpu_e2e.sh <tag>
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
for t in 3 0 6; do
  RCV_HOST_THREADS=$t timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | sed "s/^{/{\"host_threads\": $t, /" | tee -a gpurun_out/${TAG}_bench_e2e.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['host_threads'], d['value'], d['e2e']['value'], d['e2e'].get('h2d_gbs_per_gpu'))"
done
