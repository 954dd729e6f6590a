#!/bin/bash
# One GPU session: tests, memcheck, bench, reference arm, ncu launch list, ncu full capture of the vote kernel, the head and the ICP
# search kernel, evaluator stage-chain timing.
# Usage: tools/gpu_round.sh <tag> [skip-tests]
TAG=${1:-r01}
mkdir -p gpurun_out
if [ "$2" != "skip-tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.txt
  timeout 600 python -m pytest tests -m pending_gpu -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest_pending.txt   # promote to `gpu` once green
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "config1 or random_clouds or frames_api or mask_rules or large_grid or icp_vs_oracle or scene_clouds or fused or lmo_vs" tests/test_evaluator.py > gpurun_out/${TAG}_memcheck.log 2>&1
  echo "memcheck rc=$?" | tee -a gpurun_out/${TAG}_memcheck.log; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/${TAG}_memcheck.log | tail -3
fi
timeout 900 python bench.py 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/${TAG}_ref.err | tee gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --frames 512 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_vote -s 3 -c 1 -f -o gpurun_out/${TAG}_vote \
  python bench.py --steps 1 --warmup 1 --frames 256 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_vote.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_vote.log
timeout 300 python tools/head_bw.py | tee gpurun_out/${TAG}_head_bw.json
RCV_HEAD_IMAGES=48 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_head -s 3 -c 1 -f -o gpurun_out/${TAG}_head \
  python tools/head_bw.py > gpurun_out/${TAG}_ncu_head.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_head.log
timeout 300 python tools/evaluator_bw.py 2>/dev/null | tee gpurun_out/${TAG}_evaluator_bw.json
RCV_EVAL_FRAMES=64 RCV_EVAL_REPS=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_icp_corr -s 2 -c 1 -f -o gpurun_out/${TAG}_icp \
  python tools/evaluator_bw.py > gpurun_out/${TAG}_ncu_icp.log 2>&1
for v in icp4 icp1; do
  if [ -f build/librcvvote_$v.so ]; then RCV_LIB_PATH=$PWD/build/librcvvote_$v.so timeout 300 python tools/evaluator_bw.py 2>/dev/null | sed "s/^{/{\"variant\": \"$v\", /" | tee -a gpurun_out/${TAG}_evaluator_bw.json; fi
done
timeout 300 python tools/config2_bench.py 2>gpurun_out/${TAG}_config2.err | tee gpurun_out/${TAG}_config2.json; tail -2 gpurun_out/${TAG}_config2.err
ls -la gpurun_out/ | tail -20
