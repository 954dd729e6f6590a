"""Times k_head1x1 (the 1x1 head, K5) on synthetic conv7 outputs and prints one JSON line."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rcvpose_b200 import api

B, H, W = int(os.environ.get("RCV_HEAD_IMAGES", "192")), 480, 640
ctx = api.VoteContext(0, max_items=4, max_points_total=1024, max_grid=64)
up = torch.relu(torch.randn((B, 32, H, W), device="cuda")).to(torch.bfloat16)
w = torch.randn((2, 32), device="cuda")
b = torch.randn(2, device="cuda")
for _ in range(3):
    out = ctx.head_1x1(up, w, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    out = ctx.head_1x1(up, w, b)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
byt = up.numel() * 2 + out.numel() * 4
print(json.dumps({"kernel": "k_head1x1", "images": B, "ms": round(ms, 4), "GB_per_s": round(byt / ms / 1e6, 1),
                  "bytes_per_launch": byt}))
