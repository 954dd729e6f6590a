#!/bin/bash
TAG=${1:-r02n}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
tools/gpu_variants.sh $TAG
for v in rfirst t640rf; do RCV_LIB_PATH=$PWD/build/librcvvote_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config1 or random_clouds or frames_api or ycb_shaped or fine_voxel" 2>&1 | tail -2; done
