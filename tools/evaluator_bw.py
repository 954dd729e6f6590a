"""Times the evaluator's device-side stage chain (SURVEY.md 8f N1 + N3) on synthetic LINEMOD-shaped frames and prints one JSON line:
rcv_vote_frames -> rcv_horn_batch -> rcv_add_metric_batch -> rcv_scene_clouds -> rcv_icp_batch -> rcv_add_metric_batch.
CUDA events on torch's current stream (the stream every call is launched on), 1 warm-up + REPS timed repetitions."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rcvpose_b200 import api, synth

B = int(os.environ.get("RCV_EVAL_FRAMES", "256"))
M = int(os.environ.get("RCV_EVAL_CAD", "5841"))          # vertices of LINEMOD's ape.ply
REPS = int(os.environ.get("RCV_EVAL_REPS", "3"))
OBJ_R = 55.0
ctx = api.VoteContext(0, max_items=B * 3, max_points_total=B * 3 * 16384, max_grid=256)
data = synth.torch_batch(B, 3, seed=77, obj_radius_mm=(OBJ_R, OBJ_R))
K = torch.from_numpy(synth.linemod_K).cuda()
rng = np.random.default_rng(0)
u = rng.normal(size=(M, 3))
cad = torch.from_numpy(u / np.linalg.norm(u, axis=1, keepdims=True) * OBJ_R).cuda()      # object frame = sphere centre, mm
gt = torch.zeros((B, 4, 4), dtype=torch.float64, device="cuda")
gt[:, :3, :3] = torch.eye(3, dtype=torch.float64, device="cuda")
gt[:, :3, 3] = data["centre_mm"]
gt[:, 3, 3] = 1.0
stages = ["vote_frames", "horn", "add_before", "scene_clouds", "icp", "add_after"]


def chain(ev=None):
    def mark(i):
        if ev is not None:
            ev[i].record()
    mark(0)
    out = ctx.vote_frames(data["depth"], data["radius"], K, mask_flags=api.RCV_MASK_RADIUS_NONZERO)
    mark(1)
    RT = ctx.horn_batch(data["model_mm"], out["centre_mm"])
    mark(2)
    mean, mn = ctx.add_metric(cad, RT, gt)
    mark(3)
    scene, offs, _ = ctx.scene_clouds(data["depth"], data["radius"], K, mask_flags=api.RCV_MASK_RADIUS_NONZERO)
    mark(4)
    reg = ctx.icp(cad, scene, offs, RT, mean)
    mark(5)
    mean2, mn2 = ctx.add_metric(cad, reg["RT"], gt)
    mark(6)
    return out, mean, offs, reg, mean2


chain()
torch.cuda.synchronize()
tot = np.zeros(len(stages))
for _ in range(REPS):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)]
    out, mean, offs, reg, mean2 = chain(ev)
    torch.cuda.synchronize()
    tot += np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(len(stages))])
ms = tot / REPS
scene_n = (offs[1:] - offs[:-1]).double()
iters = reg["iters"].double()
pairs = float((scene_n * (iters + 1)).sum()) * M          # nearest-neighbour distance evaluations of the ICP call
print(json.dumps({
    "tool": "evaluator_bw", "frames": B, "cad_points": M, "reps": REPS,
    "stage_ms": {s: round(float(m), 3) for s, m in zip(stages, ms)}, "chain_ms": round(float(ms.sum()), 3),
    "frames_per_s": round(B / ms.sum() * 1e3, 1), "frames_per_s_after_voting": round(B / ms[1:].sum() * 1e3, 1),
    "scene_points_mean": round(float(scene_n.mean()), 1), "icp_iters_mean": round(float(iters.mean()), 2), "icp_iters_max": int(iters.max()),
    "icp_fitness_mean": round(float(reg["fitness"].mean()), 4), "icp_pair_tests": pairs, "icp_gpairs_per_s": round(pairs / ms[4] / 1e6, 1),
    "add_pairs_per_s_G": round(B * M * M / ms[2] / 1e6, 1),
    "add_before_mm_mean": round(float(mean.mean()), 4), "add_after_mm_mean": round(float(mean2.mean()), 4),
    "status_nonzero": int((out["status"] != 0).sum())}))
