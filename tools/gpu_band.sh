#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python tools/fine_voxel_bw.py 2>&1 | grep "^{" | tee gpurun_out/${TAG}_fine_voxel.jsonl
timeout 600 python bench.py --workload ycb --frames 512 --steps 3 --warmup 3 --no-cpu 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench_ycb.json | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-200
