#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_evaluator.py tests/test_ycb_evaluator.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest_eval.txt
timeout 600 python tools/evaluator_bw.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_evaluator_bw.json | cut -c1-900
RCV_ICP_BRUTE=1 timeout 600 python tools/evaluator_bw.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_evaluator_bw_brute.json | cut -c1-900
