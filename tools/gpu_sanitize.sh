#!/bin/bash
# memcheck on the kernels added late in round 2 (conv7 head, grid ICP / ADD, band tiles, host entry point) + racecheck on the vote kernel
TAG=$1
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py tests/test_evaluator.py tests/test_producer.py -m gpu -x -q \
  -k "config1 or random_clouds or frames_api or host_entry or large_grid or fine_voxel or ycb_shaped or add_metric or icp or scene_clouds or conv7_head_kernel_vs_torch or two_call" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/${TAG}_memcheck.log; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1 or random_clouds or large_grid" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/${TAG}_racecheck.log; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/${TAG}_racecheck.log | tail -3
