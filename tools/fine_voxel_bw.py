"""Throughput of the vote path on the fine-voxel stress shape (BASELINE configs[4]): D ~ 500 grids, row-band tiles."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from rcvpose_b200 import api

rng = np.random.default_rng(3)
items, n = 8, 4000
clouds, radii = [], []
for b in range(items):
    xyz = rng.normal(0, 0.02, size=(n, 3)) + np.array([0.05, -0.02, 0.85])
    kp = xyz.mean(0) + np.array([0.11, 0.06, -0.09])
    clouds.append(xyz)
    radii.append((np.linalg.norm(xyz - kp, axis=1) * 10 + rng.normal(0, 0.005, n)).astype(np.float32))
X = torch.from_numpy(np.concatenate(clouds)).cuda()
R = torch.from_numpy(np.concatenate(radii)).cuda()
off = torch.arange(0, (items + 1) * n, n, dtype=torch.int64).cuda()
for unit in (5.0, 2.5, 1.2):
    ctx = api.VoteContext(0, max_items=items, max_points_total=items * n, max_grid=640, max_units=items * 640 * 8)
    out = ctx.vote_points(X, R, off, acc_unit=unit)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = ctx.vote_points(X, R, off, acc_unit=unit)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    votes = int(out["votes"].sum().item())
    print(json.dumps({"acc_unit_mm": unit, "items": items, "points_per_item": n, "grid": int(out["grid"][0]), "status": int(out["status"].max()),
                      "votes": votes, "ms": round(ms, 3), "Gvotes_per_s": round(votes / ms / 1e6, 1)}))
