#!/bin/bash
# Head kernel check: parity test + bandwidth sweep.  Usage: tools/gpu_head.sh [tag]
TAG=${1:-head}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -k head -x -q 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.txt
rm -f gpurun_out/${TAG}_bw.txt
for cfg in "21 16" "21 8" "22 8" "24 3" "32 4" "34 2" "31 8"; do set -- $cfg
RCV_HEAD_CFG=$1 RCV_HEAD_CTAS=$2 timeout 300 python - <<'PY' 2>&1 | tee -a gpurun_out/${TAG}_bw.txt
import torch, json, os
from rcvpose_b200 import api
ctx = api.VoteContext(0, max_items=4, max_points_total=1024, max_grid=64)
B, H, W = 192, 480, 640
up = torch.relu(torch.randn((B, 32, H, W), device="cuda")).to(torch.bfloat16)
w = torch.randn((2, 32), device="cuda"); b = torch.randn(2, device="cuda")
for _ in range(3): out = ctx.head_1x1(up, w, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): out = ctx.head_1x1(up, w, b)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = up.numel() * 2 + out.numel() * 4
print(json.dumps({"kernel": "k_head1x1", "cfg": os.environ.get("RCV_HEAD_CFG"), "ctas_per_sm": os.environ.get("RCV_HEAD_CTAS"), "images": B, "ms": round(ms, 4), "GB_per_s": round(byt / ms / 1e6, 1)}))
PY
done
