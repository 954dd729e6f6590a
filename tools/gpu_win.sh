#!/bin/bash
TAG=$1
mkdir -p gpurun_out
for v in 0 1 0 1; do
  RCV_FULL_WINDOW=$v timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | sed "s/^{/{\"full_window\": $v, /" | tee -a gpurun_out/${TAG}_bench_variants.json | cut -c1-120
done
RCV_FULL_WINDOW=0 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_op_shared_atom.sum --clock-control none -k regex:k_vote_runs -s 2 -c 1 --csv --log-file gpurun_out/${TAG}_win0.csv python bench.py --steps 1 --warmup 1 --frames 512 --no-e2e --no-cpu > /dev/null 2>&1
RCV_FULL_WINDOW=1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__inst_executed_op_shared_atom.sum --clock-control none -k regex:k_vote_runs -s 2 -c 1 --csv --log-file gpurun_out/${TAG}_win1.csv python bench.py --steps 1 --warmup 1 --frames 512 --no-e2e --no-cpu > /dev/null 2>&1
tail -3 gpurun_out/${TAG}_win0.csv | cut -c150-330; tail -3 gpurun_out/${TAG}_win1.csv | cut -c150-330
