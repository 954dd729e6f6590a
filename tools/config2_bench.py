"""BASELINE configs[1] (LINEMOD 'ape' full test path on synthetic RGB-D): RGB -> three FCN-ResNet trunks (PyTorch bf16, random
weights: there are no checkpoints offline) -> fused tcgen05 head + mask rule + vote -> Horn pose, on one B200.  Prints one JSON line
with the time per stage (CUDA events on torch's current stream).  Three producer variants: eager, CUDA-graphed, CUDA-graphed with channels_last trunks.  Because untrained networks give meaningless radii, the keypoints are not checked here --
stage-wise parity is in tests/test_producer.py (trunk vs the reference model) and tests/test_evaluator.py (fused head + vote)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rcvpose_b200 import api, producer, synth

B = int(os.environ.get("RCV_C2_FRAMES", "8"))
REPS = int(os.environ.get("RCV_C2_REPS", "3"))
ctx = api.VoteContext(0, max_items=B * 3, max_points_total=B * 3 * 65536, max_grid=400)
torch.manual_seed(0)
stage = producer.ProducerStage([producer.RadiusTrunk() for _ in range(3)], ctx)
frames = [synth.config3_frame(100 + f) for f in range(B)]
depth = torch.from_numpy(np.stack([f["depth"] for f in frames]).view(np.int16)).cuda()
rgb = producer.normalise_rgb(np.random.default_rng(0).integers(0, 256, size=(B, 480, 640, 3)).astype(np.uint8)).cuda()
K = torch.from_numpy(synth.linemod_K).cuda()
max_radii = torch.full((3,), 2.5, dtype=torch.float64, device="cuda")                       # dm; 'ape'-sized object
model_mm = torch.from_numpy(np.stack([f["kpts_mm"] - f["centre_mm"] for f in frames])).cuda().contiguous()
flags = api.RCV_MASK_MAX_RADIUS | api.RCV_MASK_RADIUS_POSITIVE    # random-init seg scores never reach 0.8: keep the radius rules only


def measure(stage, label):
    def step(ev=None):
        mark = (lambda i: ev[i].record()) if ev is not None else (lambda i: None)
        mark(0)
        up = stage.activations(rgb)
        mark(1)
        if stage.fuse_tail:
            out = ctx.conv7_head_vote_frames(up, *stage.tail_params, depth, K, max_radii=max_radii, mask_flags=flags)
        else:
            out = ctx.head_vote_frames(up, stage.weight, stage.bias, depth, K, max_radii=max_radii, mask_flags=flags)
        mark(2)
        RT = ctx.horn_batch(model_mm, out["centre_mm"])
        mark(3)
        return out, RT

    step()
    torch.cuda.synchronize()
    tot = np.zeros(3)
    for _ in range(REPS):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        out, RT = step(ev)
        torch.cuda.synchronize()
        tot += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(3)])
    ms = tot / REPS
    return out, {"producer": label, "stage_ms": {"trunks_pytorch_bf16": round(float(ms[0]), 3), "head_mask_vote": round(float(ms[1]), 3),
                                                 "horn": round(float(ms[2]), 3)},
                 "frames_per_s": round(B / ms.sum() * 1e3, 2), "frames_per_s_after_trunks": round(B / ms[1:].sum() * 1e3, 1),
                 "trunk_tflops": round(3 * 265.6e9 * B / (ms[0] * 1e-3) / 1e12, 1)}


rows = []
out, r = measure(stage, "eager, one trunk after the other")
rows.append(r)
ref_centres = out["centre_mm"].clone()
out, r = measure(stage.capture(B, 480, 640, concurrent=True), "one CUDA graph, the three trunks on forked streams")
rows.append(r)
same = bool(torch.equal(out["centre_mm"], ref_centres))
del stage
torch.manual_seed(0)
stage_cl = producer.ProducerStage([producer.RadiusTrunk() for _ in range(3)], ctx, channels_last=True).capture(B, 480, 640, concurrent=True)
out, r = measure(stage_cl, "CUDA graph + channels_last trunks")
rows.append(r)
del stage_cl
torch.manual_seed(0)
stage_ft = producer.ProducerStage([producer.RadiusTrunk() for _ in range(3)], ctx, fuse_tail=True).capture(B, 480, 640, concurrent=True)
out, r = measure(stage_ft, "CUDA graph + channels_last trunks up to up1 + fused tail kernel (conv7+BN+ReLU+conv8+mask rule)")
rows.append(r)
print(json.dumps({"tool": "config2_bench", "workload": "BASELINE configs[1]: RGB -> 3 x FCN-ResNet-152 trunk (bf16, random weights) -> fused head + vote -> Horn",
                  "frames": B, "reps": REPS, "variants": rows, "graph_output_equals_eager": same,
                  "points_per_item_mean": round(float(out["n_points"].float().mean()), 1), "status_nonzero": int((out["status"] != 0).sum())}))
