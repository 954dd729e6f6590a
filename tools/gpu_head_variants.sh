#!/bin/bash
# Bandwidth of k_head1x1 for library variants built by tools/build_variant.sh.  Usage: tools/gpu_head_variants.sh <tag> <variant...>
TAG=$1; shift
mkdir -p gpurun_out; rm -f gpurun_out/${TAG}_bw.txt
for v in default "$@"; do
  if [ "$v" = default ]; then unset RCV_LIB_PATH; else export RCV_LIB_PATH=$PWD/build/librcvvote_$v.so; fi
  timeout 300 python -m pytest tests/test_gpu_parity.py -k head -x -q 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_bw.txt
  for cfg in "21 16" "21 12" "22 8" "24 3" "24 4"; do set -- $cfg
    echo -n "$v " | tee -a gpurun_out/${TAG}_bw.txt
    RCV_HEAD_CFG=$1 RCV_HEAD_CTAS=$2 timeout 300 python tools/head_bw.py 2>&1 | tail -1 | sed "s/^{/{\"cfg\": \"$1x$2\", /" | tee -a gpurun_out/${TAG}_bw.txt
  done
done
