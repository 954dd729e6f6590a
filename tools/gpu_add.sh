#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_evaluator.py tests/test_ycb_evaluator.py -m gpu -x -q -k "add or evaluator or estimate or icp or scene" 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_add.txt
timeout 600 python tools/evaluator_bw.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_evaluator_bw.json | cut -c1-700
