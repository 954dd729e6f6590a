"""The producer's tail (conv7 + BN + ReLU + conv8) on B x 480 x 640 maps: PyTorch bf16 channels_last conv7 module + rcv_head_1x1
against the one-kernel rcv_conv7_head.  Prints one JSON line (CUDA events, median of REPS)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rcvpose_b200 import api, producer

B = int(os.environ.get("RCV_TAIL_IMAGES", "24"))
REPS = int(os.environ.get("RCV_TAIL_REPS", "20"))
H, W = 480, 640
ctx = api.VoteContext(0, max_items=8, max_points_total=1 << 20, max_grid=256)
torch.manual_seed(0)
t = producer.RadiusTrunk()
conv7 = t.conv7.to(device="cuda", dtype=torch.bfloat16).eval().to(memory_format=torch.channels_last)
w8, b8 = [a.cuda() for a in t.head()]
t = t.to(device="cuda", dtype=torch.bfloat16)
w7, sc, sh, _, _ = [a.cuda() for a in t.tail()]
x = (torch.randn((B, 64, H, W), device="cuda") * 0.5).bfloat16().contiguous(memory_format=torch.channels_last)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(REPS):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


with torch.no_grad():
    def unfused():
        up = conv7(x).contiguous()            # NCHW planes: what the head's TMA boxes read
        return ctx.head_1x1(up, w8, b8)

    def unfused_conv_only():
        return conv7(x)

    def fused():
        return ctx.conv7_head(x, w7, sc, sh, w8, b8)

    a = unfused(); b = fused()
    torch.cuda.synchronize()
    err = float((a - b).abs().max()); ref = float(a.abs().max())
    ms_u, ms_c, ms_f = timed(unfused), timed(unfused_conv_only), timed(fused)
px = B * H * W
print(json.dumps({"tool": "tail_bench", "images": B, "shape": [H, W], "ms_unfused_conv7_bn_relu_plus_head": round(ms_u, 3),
                  "ms_unfused_conv7_bn_relu_only": round(ms_c, 3), "ms_fused_kernel": round(ms_f, 3),
                  "fused_us_per_image": round(ms_f * 1e3 / B, 1), "fused_tflops": round(px * 2 * 64 * 9 * 32 / (ms_f * 1e-3) / 1e12, 1),
                  "fused_hbm_gbs_algorithmic": round(px * (128 + 8) / (ms_f * 1e-3) / 1e9, 0),
                  "max_abs_diff_fused_vs_unfused": err, "max_abs_value": ref}))
