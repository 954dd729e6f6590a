#!/bin/bash
# Round-end style check of the default build: tools/gpu_final.sh <tag>
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench_reference.json | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vote -s 3 -c 1 -f -o gpurun_out/${TAG}_vote \
  python bench.py --steps 1 --warmup 1 --frames 256 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_vote.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_vote.log
# dram traffic of the full-size launch (4096 frames): two metrics only, one replay pass
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_vote -s 2 -c 1 --csv \
  --log-file gpurun_out/${TAG}_vote_traffic_full.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
tail -4 gpurun_out/${TAG}_vote_traffic_full.csv
