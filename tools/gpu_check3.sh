#!/bin/bash
# Tests (all GPU tests incl. the new producer / evaluator routes), memcheck + racecheck on the vote kernel, config2 variants, bench.
TAG=${1:-r02m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python tools/config2_bench.py 2>gpurun_out/${TAG}_config2.err | tee gpurun_out/${TAG}_config2.json; tail -3 gpurun_out/${TAG}_config2.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py tests/test_ycb_evaluator.py -m gpu -x -q \
  -k "config1 or random_clouds or frames_api or mask_rules or ycb_shaped or fine_voxel or estimate_6d_pose_ycb" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/${TAG}_memcheck.log; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log | tail -3
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1 or random_clouds" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck rc=$?" | tee -a gpurun_out/${TAG}_racecheck.log; grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/${TAG}_racecheck.log | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-200
