#!/bin/bash
TAG=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 | tee gpurun_out/${TAG}_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:k_frame_mask -s 2 -c 2 --csv --log-file gpurun_out/${TAG}_k1.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > /dev/null 2>&1
tail -4 gpurun_out/${TAG}_k1.csv | cut -c100-400
