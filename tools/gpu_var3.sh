#!/bin/bash
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.txt
tools/gpu_variants.sh $TAG
