#!/bin/bash
# First GPU call of round 2: everything that has never run on hardware.
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python -m pytest tests -m pending_gpu -q 2>&1 | tail -40 | tee gpurun_out/${TAG}_pytest_pending.txt
timeout 600 python tools/config2_bench.py 2>gpurun_out/${TAG}_config2.err | tee gpurun_out/${TAG}_config2.json; tail -20 gpurun_out/${TAG}_config2.err
timeout 600 python bench.py 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
