"""Batched evaluator for the voting path: frames x keypoints -> keypoint centres -> Horn poses,
sharded over the GPUs of one node.

Replaces the per-image Python loop body of the reference's estimate_6d_pose_lm
(AccumulatorSpace.py:553-662: mask, rgbd_to_point_cloud, Accumulator_3D per keypoint, then
horn.lmshorn per image).  Frames and keypoints are independent, so ranks own contiguous frame ranges
and run with no data-path collective; the only communication is one small all_gather of the
per-frame results (about 250 bytes per frame).
"""
import torch
import torch.distributed as dist

from . import api


def shard_range(n_frames, rank, world):
    """Contiguous frame range [lo, hi) of `rank`: the first n % world ranks take one extra frame."""
    q, r = divmod(int(n_frames), int(world))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_by_cost(costs, world):
    """Contiguous frame ranges with (nearly) equal summed cost -- SURVEY.md 8e: "for YCB-shaped work, balance by sum N*R^2
    rather than by frame count, since per-item cost varies > 10x".  Rank r ends at the first frame where the running cost
    reaches (r + 1) / world of the total.  Returns [(lo, hi)] * world; every frame belongs to exactly one range."""
    import numpy as np
    c = np.asarray(costs, dtype=np.float64)
    n = len(c)
    cum = np.cumsum(c)
    total = float(cum[-1]) if n else 0.0
    bounds = [0]
    for r in range(1, int(world)):
        t = total * r / world
        k = int(np.searchsorted(cum, t, side="left")) + 1 if total > 0 else (n * r) // world
        # the frame that crosses the target goes to whichever side leaves the smaller error
        if total > 0 and 0 < k <= n and abs(cum[k - 2] - t if k >= 2 else t) < abs(cum[k - 1] - t):
            k -= 1
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(int(world))]


def pack_results(centres_mm, RT, peak, status):
    """(B,Kp,3) f64, (B,4,4) f64, (B,Kp) i32, (B,Kp) i32 -> one (B, 3Kp+16+2Kp) f64 row per frame."""
    B = centres_mm.shape[0]
    return torch.cat([centres_mm.reshape(B, -1), RT.reshape(B, 16), peak.reshape(B, -1).to(torch.float64),
                      status.reshape(B, -1).to(torch.float64)], dim=1).contiguous()


def unpack_results(rows, n_kpts):
    B = rows.shape[0]
    o = 0
    centres = rows[:, o:o + 3 * n_kpts].reshape(B, n_kpts, 3); o += 3 * n_kpts
    RT = rows[:, o:o + 16].reshape(B, 4, 4); o += 16
    peak = rows[:, o:o + n_kpts].to(torch.int32); o += n_kpts
    status = rows[:, o:o + n_kpts].to(torch.int32)
    return centres, RT, peak, status


def gather_results(rows, counts=None, group=None):
    """All-gather the per-rank result rows (ranks may own different frame counts).  Works on CUDA
    tensors over NCCL and on CPU tensors over gloo (used by the CPU tests of the sharding logic)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    if counts is None:
        n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
        allc = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(allc, n, group=group)
        counts = [int(c.item()) for c in allc]
    m = max(counts)
    pad = rows if rows.shape[0] == m else torch.cat([rows, rows.new_zeros((m - rows.shape[0], rows.shape[1]))], dim=0)
    out = rows.new_empty((world * m, rows.shape[1]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return torch.cat([out[r * m: r * m + counts[r]] for r in range(world)], dim=0)


class VotingPipeline:
    """One per process/GPU.  step() = K1 (mask + back-projection + compaction) -> prelude -> K2 vote
    (+ fused peak) -> finalize -> K4 Horn, all asynchronous on the current stream."""

    def __init__(self, device=0, max_frames=4096, n_kpts=3, max_points_total=1 << 27, max_grid=256, image=(480, 640), max_model_points=0):
        # image / max_model_points: the context allocates its image- and model-sized scratch now, not on the first hot call
        self.ctx = api.VoteContext(device, max_items=max_frames * n_kpts, max_points_total=max_points_total, max_grid=max_grid, image=image,
                                   max_model_points=max_model_points)
        self.n_kpts = n_kpts

    def step(self, depth, radius, K, model_mm, sem=None, max_radii=None, mask_flags=api.RCV_MASK_RADIUS_NONZERO, **kw):
        out = self.ctx.vote_frames(depth, radius, K, sem=sem, max_radii=max_radii, mask_flags=mask_flags, **kw)
        out["RT"] = self.ctx.horn_batch(model_mm, out["centre_mm"])
        return out

    def add_pass(self, cad_mm, RT_est, RT_gt, threshold_mm, symmetric=False):
        """ADD(-S) bookkeeping of the reference's evaluator before ICP (AccumulatorSpace.py:687-695): per frame, does the mean
        (or, for a symmetric class, the minimum) nearest-neighbour distance between the CAD points under the ground-truth and
        under the estimated pose stay within `threshold_mm` (the reference uses add_threshold[class] * 1000)?  Returns
        (passed (B,) bool, distance (B,) float64)."""
        mean, mn = self.ctx.add_metric(cad_mm, RT_est, RT_gt)
        d = mn if symmetric else mean
        return d <= threshold_mm, d

    def step_gathered(self, depth, radius, K, model_mm, counts=None, **kw):
        out = self.step(depth, radius, K, model_mm, **kw)
        rows = gather_results(pack_results(out["centre_mm"], out["RT"], out["peak"], out["status"]), counts)
        out["gathered"] = rows
        return out
