#!/bin/bash
TAG=${1:-r02c}
mkdir -p gpurun_out
./build/ubench_atoms2 | tee gpurun_out/${TAG}_ubench_atoms2.jsonl
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ycb_shaped or fine_voxel" > gpurun_out/${TAG}_memcheck.log 2>&1
grep -E "Invalid|at 0x|by thread|Address|ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log | head -30
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
