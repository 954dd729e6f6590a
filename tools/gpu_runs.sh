#!/bin/bash
# GPU session for the run-length vote kernel: parity first, then bench of thread-count variants, then ncu.
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
for v in t768; do
  if [ -f build/librcvvote_$v.so ]; then RCV_LIB_PATH=$PWD/build/librcvvote_$v.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | sed "s/^{/{\"variant\": \"$v\", /" | tee -a gpurun_out/${TAG}_bench_variants.json; fi
done
RCV_VOTE_GEN=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | sed "s/^{/{\"variant\": \"gen1\", /" | tee -a gpurun_out/${TAG}_bench_variants.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vote -s 3 -c 1 -f -o gpurun_out/${TAG}_vote \
  python bench.py --steps 1 --warmup 1 --frames 256 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_vote.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_vote.log
