#!/bin/bash
# Vote-kernel tuning session: parity, bench, row-stride residue sweep, ncu.
TAG=${1:-r02g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
for m in 1 5 7 9 13 17 21 25 27; do
  RCV_DP_MOD=$m timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --frames 2048 2>/dev/null | sed "s/^{/{\"variant\": \"dpmod$m\", /" | tee -a gpurun_out/${TAG}_bench_variants.json | cut -c1-120
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vote -s 3 -c 1 -f -o gpurun_out/${TAG}_vote \
  python bench.py --steps 1 --warmup 1 --frames 256 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_vote.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_vote.log
