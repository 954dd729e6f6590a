#!/bin/bash
TAG=${1:-r02j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-300
