#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
