#!/bin/bash
# GPU session for the rows around the vote kernel: e2e with row cropping, large grids, YCB-shaped bench, evaluator chain,
# selected-metric ncu captures of every other kernel.
TAG=${1:-r02h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-400
timeout 300 python tools/fine_voxel_bw.py 2>/dev/null | tee gpurun_out/${TAG}_fine_voxel.jsonl
timeout 600 python bench.py --workload ycb --frames 512 --steps 3 --warmup 3 --no-cpu 2>/dev/null | tee gpurun_out/${TAG}_bench_ycb.json | cut -c1-300
timeout 300 python tools/evaluator_bw.py 2>/dev/null | tee gpurun_out/${TAG}_evaluator_bw.json
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size"
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_ncu_other_kernels.csv \
  -k regex:'k_frame_mask|k_frame_emit|k_points_from_pixels|k_scan_items|k_prelude|k_finalize|k_horn' -c 28 \
  python bench.py --steps 1 --warmup 1 --frames 1024 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_other.log 2>&1
RCV_EVAL_FRAMES=64 RCV_EVAL_REPS=1 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_ncu_evaluator_kernels.csv \
  -k regex:'k_add_nn|k_add_finish|k_icp_corr|k_icp_update|k_scene_mask|k_scene_scan|k_scene_points|k_argmax_volume|k_backproject' -c 40 \
  python tools/evaluator_bw.py > gpurun_out/${TAG}_ncu_eval.log 2>&1
ls -la gpurun_out | tail -8
