#!/bin/bash
# selected-metric ncu captures of every kernel beside the vote kernel: tools/gpu_otherk.sh <tag>
TAG=$1
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size"
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_ncu_path_kernels.csv \
  -k regex:'k_frame_mask|k_frame_emit|k_points_from_pixels|k_scan_items|k_prelude|k_finalize|k_horn' -c 28 \
  python bench.py --steps 1 --warmup 1 --frames 1024 --no-e2e --no-cpu > gpurun_out/${TAG}_ncu_other.log 2>&1
RCV_EVAL_FRAMES=64 RCV_EVAL_REPS=1 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_ncu_evaluator_kernels.csv \
  -k regex:'k_add_|k_icp_|k_scene_' -c 120 \
  python tools/evaluator_bw.py > gpurun_out/${TAG}_ncu_eval.log 2>&1
RCV_TAIL_IMAGES=24 RCV_TAIL_REPS=1 timeout 600 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${TAG}_ncu_tail_kernels.csv \
  -k regex:'k_conv7_head|k_head1x1' -c 12 \
  python tools/tail_bench.py > gpurun_out/${TAG}_ncu_tail.log 2>&1
python tools/ncu_table.py gpurun_out/${TAG}_ncu_path_kernels.csv gpurun_out/${TAG}_ncu_evaluator_kernels.csv gpurun_out/${TAG}_ncu_tail_kernels.csv | tee gpurun_out/${TAG}_other_kernels_ncu.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.txt
