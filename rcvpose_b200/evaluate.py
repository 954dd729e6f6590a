"""Batched evaluator: the caller of the voting path with the step after it (N3) and the LINEMOD
directory layout in front of it (N4).

Replaces the body of the reference's `estimate_6d_pose_lm` (AccumulatorSpace.py:495-744), which walks the test images one
by one and, per image, runs mask rule -> rgbd_to_point_cloud -> Accumulator_3D for three keypoints (:578-656), lmshorn
(:660-662), the ADD(-S) distance of the CAD model before ICP (:664-702), open3d's point-to-point ICP against the union of
the masked clouds (:620-625, :704-718) and the ADD(-S) distance after it (:723-731).  Here a batch of frames goes through
the same stages as device-resident arrays, every stage one call into librcvvote.so (include/rcvvote.h):

    rcv_vote_frames -> rcv_horn_batch -> rcv_add_metric_batch -> rcv_scene_clouds -> rcv_icp_batch -> rcv_add_metric_batch

Nothing is computed on the CPU except file reading and the pass / fail counters.
"""
import os

import numpy as np
import torch

from . import api, formats

# ---- data of the reference's evaluator (AccumulatorSpace.py:18-19, :43, :45-58, :59-61) ----
lmo_cls_names = ['ape', 'can', 'cat', 'duck', 'driller', 'eggbox', 'glue', 'holepuncher']
lm_cls_names = ['ape', 'benchvise', 'cam', 'can', 'cat', 'duck', 'driller', 'eggbox', 'glue', 'holepuncher', 'iron', 'lamp', 'phone']
lm_syms = ['eggbox', 'glue']
add_threshold = {   # 10 % of the model diameter, metres
    'eggbox': 0.019735770122546523, 'ape': 0.01421240983190395, 'cat': 0.018594838977253875, 'cam': 0.02222763033276377,
    'duck': 0.015569664208967385, 'glue': 0.01930723067998101, 'can': 0.028415044264086586, 'driller': 0.031877906042,
    'holepuncher': 0.019606109985, 'benchvise': .033091264970068, 'iron': .03172344425531, 'lamp': .03165980764376,
    'phone': .02543407135792}
linemod_K = np.array([[572.4114, 0., 325.2611], [0., 573.57043, 242.04899], [0., 0., 1.]])


def max_radii_dm(cad_m, keypoints_m, n_kpts=3):
    """Largest distance from keypoint k to a CAD point, in decimetres (AccumulatorSpace.py:539-545)."""
    out = np.zeros(n_kpts)
    for i in range(n_kpts):
        d = ((cad_m[:, 0] - keypoints_m[i + 1, 0]) ** 2 + (cad_m[:, 1] - keypoints_m[i + 1, 1]) ** 2
             + (cad_m[:, 2] - keypoints_m[i + 1, 2]) ** 2) ** 0.5
        out[i] = d.max() * 10
    return out


def _world():
    """(rank, world) of the evaluation: one process per GPU under torchrun, else (0, 1)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_frames(stems):
    """Frames are independent: rank r evaluates a contiguous range of the class's frames."""
    from .pipeline import shard_range
    rank, world = _world()
    lo, hi = shard_range(len(stems), rank, world)
    return stems[lo:hi]


def reduce_counts(values):
    """Sum of small integer counters over the ranks (the only communication of a sharded evaluation): one all_reduce of
    len(values) int64, on the GPU over NCCL or on the host over gloo."""
    import torch.distributed as dist
    rank, world = _world()
    if world == 1:
        return [int(v) for v in values]
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(v) for v in values], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(v) for v in t.tolist()]


def sync_failure(exc):
    """In a sharded evaluation a rank that fails inside its batch loop must not leave the others waiting in the final
    all_reduce: every rank reports a failure flag first, and all raise together (the failing rank its own exception)."""
    import torch.distributed as dist
    rank, world = _world()
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([1 if exc is not None else 0, rank if exc is not None else world], dtype=torch.int64, device=dev)
        flag = t[:1].clone(); who = t[1:].clone()
        dist.all_reduce(flag, op=dist.ReduceOp.SUM)
        dist.all_reduce(who, op=dist.ReduceOp.MIN)
        if exc is None and int(flag.item()) > 0:
            raise RuntimeError("evaluation failed on rank %d (%d rank(s) in all); see that rank's exception" % (int(who.item()), int(flag.item())))
    if exc is not None:
        raise exc


def _default_device(opts):
    """opts.device if given, else the process's GPU under torchrun (LOCAL_RANK), else 0."""
    d = getattr(opts, "device", None)
    return int(d) if d is not None else int(os.environ.get("LOCAL_RANK", "0"))


def _to_device(a, dev, dtype=None):
    if a is None:
        return None
    t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
    t = t.to(dev)
    if dtype is not None:
        t = t.to(dtype)
    return t.contiguous()


class FrameEvaluator:
    """Device-side stage chain for batches of frames of ONE object class (one CAD model, one keypoint set)."""

    def __init__(self, cad_mm, kpts_mm, symmetric, threshold_mm, device=0, frames_per_batch=64, image=(480, 640), n_kpts=3, max_grid=256,
                 icp=True, icp_max_iter=30, ctx=None):
        self.B, self.n_kpts, self.image = int(frames_per_batch), int(n_kpts), tuple(image)
        if ctx is None:
            ctx = api.VoteContext(device, max_items=self.B * n_kpts, max_points_total=self.B * n_kpts * image[0] * image[1], max_grid=max_grid,
                                  image=self.image, max_model_points=len(cad_mm))
        self.ctx = ctx                     # (a ProducerStage's own context when the maps come from the fused head)
        self.dev = ctx.device
        self.cad_mm = torch.as_tensor(np.ascontiguousarray(cad_mm, dtype=np.float64), device=self.dev)
        self.kpts_mm = torch.as_tensor(np.ascontiguousarray(kpts_mm, dtype=np.float64), device=self.dev)
        self.symmetric, self.threshold_mm = bool(symmetric), float(threshold_mm)
        self.icp, self.icp_max_iter = bool(icp), int(icp_max_iter)

    def _gt(self, RT_gt_mm, B):
        gt = torch.zeros((B, 4, 4), dtype=torch.float64, device=self.dev)
        g = _to_device(RT_gt_mm, self.dev, torch.float64)
        gt[:, :g.shape[1], :] = g
        gt[:, 3, 3] = 1.0
        return gt

    def _after_vote(self, out, gt, scene_fn, centres_override=None, icp_rel_fitness=1e-6, icp_rel_rmse=1e-6, zero_empty_centres=False,
                    icp_threshold_mean=False):
        """Horn -> ADD(-S) -> [scene cloud -> ICP -> ADD(-S)] on the keypoint centres of a frames call."""
        ctx, dev = self.ctx, self.dev
        centres = out["centre_mm"]
        if zero_empty_centres:   # the LMO evaluator leaves the row of a keypoint without pixels at zero (AccumulatorSpace.py:795, :858)
            centres[(out["status"] & api.RCV_ST_EMPTY_MASK) != 0] = 0.0
        if centres_override is not None:   # frames voted outside the fused path (float64 radius maps), see LinemodEvaluator
            idx, val = centres_override
            centres[idx[0].to(dev), idx[1].to(dev)] = _to_device(val, dev, torch.float64)
        RT = ctx.horn_batch(self.kpts_mm, centres)
        mean, mn = ctx.add_metric(self.cad_mm, RT, gt)
        before = mn if self.symmetric else mean
        res = dict(centre_mm=centres, RT=RT, dist_before=before, passed_before=before <= self.threshold_mm, status=out["status"],
                   n_points=out["n_points"], peak=out["peak"], grid=out["grid"], mean_before=mean)
        if self.icp:
            scene, offs, _ = scene_fn()
            # correspondence threshold: the ADD(-S) figure of the class (LM / LMO, :706-707, :929) or always the mean distance (YCB, :1151)
            thr = (mean if icp_threshold_mean else before).contiguous()
            reg = ctx.icp(self.cad_mm, scene, offs, RT, thr, max_iter=self.icp_max_iter, rel_fitness=icp_rel_fitness, rel_rmse=icp_rel_rmse)
            mean2, mn2 = ctx.add_metric(self.cad_mm, reg["RT"], gt)
            after = mn2 if self.symmetric else mean2
            res.update(RT_icp=reg["RT"], dist_after=after, passed_after=after <= self.threshold_mm, icp_fitness=reg["fitness"],
                       icp_rmse=reg["rmse"], icp_iters=reg["iters"], scene_points=offs[1:] - offs[:-1])
        return {k: v.cpu().numpy() for k, v in res.items()}

    def run(self, depth, radius, K, RT_gt_mm, max_radii=None, sem=None, mask_flags=api.MASK_LM_NPY, sem_threshold=0.8, depth_div=1.0,
            scene_scale=1.0, centres_override=None, icp_rel_fitness=1e-6, icp_rel_rmse=1e-6, zero_empty_centres=False,
            icp_threshold_mean=False, **vote_kw):
        """depth (B,H,W), radius (B,Kp,H,W) float32 [, sem (B,Kp,H,W) float32], K (3,3) or (B,3,3), RT_gt_mm (B,3,4) or (B,4,4)
        (translation in mm) -- NumPy or torch, host or device.  Returns a dict of HOST arrays: centre_mm (B,Kp,3), RT (B,4,4),
        dist_before (B,), RT_icp (B,4,4), dist_after (B,), passed_before / passed_after (B,) bool, status (B,Kp), n_points (B,Kp),
        icp_fitness / icp_rmse / icp_iters, scene_points (B,)."""
        dev, ctx = self.dev, self.ctx
        to = lambda a, dt=None: _to_device(a, dev, dt)  # noqa: E731
        depth, radius, sem = to(depth), to(radius, torch.float32), to(sem, torch.float32)
        K, max_radii = to(K, torch.float64), to(max_radii, torch.float64)
        gt = self._gt(RT_gt_mm, radius.shape[0])
        out = ctx.vote_frames(depth, radius, K, sem=sem, max_radii=max_radii, mask_flags=mask_flags, sem_threshold=sem_threshold,
                              depth_div=depth_div, **vote_kw)
        scene_fn = lambda: ctx.scene_clouds(depth, radius, K, sem=sem, max_radii=max_radii, mask_flags=mask_flags, sem_threshold=sem_threshold,  # noqa: E731
                                            depth_div=depth_div, scale=scene_scale)
        return self._after_vote(out, gt, scene_fn, centres_override, icp_rel_fitness, icp_rel_rmse, zero_empty_centres, icp_threshold_mean)

    def run_stage(self, stage, images, depth, K, RT_gt_mm, max_radii, mask_flags=api.MASK_LM_CKPT, sem_threshold=0.8, depth_div=1.0,
                  scene_scale=1.0, icp_rel_fitness=1e-6, icp_rel_rmse=1e-6, zero_empty_centres=False, icp_threshold_mean=False, **vote_kw):
        """The checkpoint branch through the fused producer: images (B,3,H,W) normalised RGB -> the stage's three
        trunks -> rcv_head_vote_frames (conv8 + mask rule + vote; the seg plane never reaches memory) -> the same chain as run().
        The ICP scene cloud comes from the survival bits that call left behind (rcv_scene_clouds_last).  `stage` must have been
        built on this evaluator's context."""
        assert stage.ctx is self.ctx, "build the FrameEvaluator with ctx=stage.ctx"
        dev = self.dev
        depth, K, max_radii = _to_device(depth, dev), _to_device(K, dev, torch.float64), _to_device(max_radii, dev, torch.float64)
        gt = self._gt(RT_gt_mm, depth.shape[0])
        out = stage.vote(images, depth, K, max_radii, mask_flags=mask_flags, sem_threshold=sem_threshold, depth_div=depth_div, **vote_kw)
        scene_fn = lambda: self.ctx.scene_clouds_last(depth, K, self.n_kpts, depth_div=depth_div, scale=scene_scale)  # noqa: E731
        return self._after_vote(out, gt, scene_fn, None, icp_rel_fitness, icp_rel_rmse, zero_empty_centres, icp_threshold_mean)


def _revote_oversized(res, status_mask, radius, sem, depth, K, rule, shim, override_idx, override_val):
    """Items whose accumulator is larger than the evaluator context allows (RCV_ST_D_EXCEEDS_CAP: one false-positive mask pixel on
    background depth is enough) are voted again through the exact drop-in surface, whose context takes grids up to 640 -- the
    reference handles any size, slowly.  rule(radius_map, sem_map) -> boolean mask of surviving pixels (before depth != 0).
    Appends to the override lists; returns True if anything was re-voted."""
    todo = np.argwhere(status_mask)
    for i, k in todo:
        m = rule(radius[i, k], None if sem is None else sem[i, k], k)
        dm = depth[i] * np.where(m, 1, 0)
        xyz = shim.rgbd_to_point_cloud(K, dm)
        override_idx.append((int(i), int(k)))
        override_val.append(shim.Accumulator_3D(xyz / 1000, radius[i, k][dm.nonzero()])[0])
    return len(todo) > 0


class LinemodClass:
    """One class of the reference's LINEMOD layout (AccumulatorSpace.py:500-551, :566, :599, :612):
        <root>/LINEMOD/<cls>/{<cls>.ply, Outside9.npy, Split/val.txt, JPEGImages/<stem>.jpg, pose/pose<N>.npy}
        <root>/LINEMOD_ORIG/<cls>/data/depth<N>.dpt
        <root>/LINEMOD_ORIG/estRadialMap/<cls>/Out_pt<k>_dm/<stem>.npy        (k = 1..3, decimetres)"""

    def __init__(self, root_dataset, class_name):
        self.name = class_name
        self.pv = root_dataset + "LINEMOD/" + class_name + "/"
        self.orig = root_dataset + "LINEMOD_ORIG/" + class_name + "/"
        self.est = os.path.join(root_dataset + "LINEMOD_ORIG/", "estRadialMap", class_name)
        self.cad_m = formats.read_ply_points(self.pv + class_name + ".ply")
        self.keypoints_m = formats.load_keypoints(self.pv + "Outside9.npy")
        self.max_radii_dm = max_radii_dm(self.cad_m, self.keypoints_m)
        test = set(formats.read_split(self.pv + "Split/val.txt"))
        self.stems = sorted(os.path.splitext(f)[0] for f in os.listdir(self.pv + "JPEGImages/")
                            if f.endswith(".jpg") and os.path.splitext(f)[0] in test)

    def image_path(self, stem):
        return self.pv + "JPEGImages/" + stem + ".jpg"

    def depth(self, stem):
        return formats.read_depth(self.orig + "data/depth" + str(int(stem)) + ".dpt")

    def pose_mm(self, stem):
        rt = formats.load_pose(self.pv + "pose/pose" + str(int(stem)) + ".npy").copy()
        rt[:, 3] = rt[:, 3] * 1000
        return rt

    def radial_est(self, stem, k):
        return np.load(os.path.join(self.est, "Out_pt" + str(k) + "_dm", stem + ".npy"))


def evaluate_lm_class(root_dataset, class_name, using_ckpts=False, producer=None, device=0, frames_per_batch=64, icp=True, verbose=True,
                      max_grid=256):
    """One class of estimate_6d_pose_lm.  `producer(class_name, k, image_path) -> (sem, radial)` supplies the (H,W) float32 maps of
    keypoint k = 1..3 when using_ckpts (the reference's FCResBackbone + DenseFCNResNet152 stay PyTorch code outside this package).
    Returns dict(n, add_before, add_after, frames=[...], per-frame arrays)."""
    from . import AccumulatorSpace as shim
    cls = LinemodClass(root_dataset, class_name)
    sym = class_name in lm_syms
    if using_ckpts and producer is None:
        raise ValueError("using_ckpts needs a producer(class_name, k, image_path) -> (sem, radial), or a rcvpose_b200.producer.ProducerStage "
                         "holding the class's three networks")
    staged = using_ckpts and hasattr(producer, "vote") and hasattr(producer, "activations")     # a ProducerStage: fused head + vote (N2)
    ev = None
    acc = {}
    my_stems = shard_frames(cls.stems)        # under torchrun every rank takes a contiguous range of the frames
    verbose = verbose and _world()[0] == 0
    failure = None
    try:
        for b0 in range(0, len(my_stems), frames_per_batch):
            stems = my_stems[b0:b0 + frames_per_batch]
            depth = np.stack([cls.depth(s) for s in stems])
            H, W = depth.shape[1:]
            if staged:
                # RGB -> three trunks -> fused conv8 + mask rule (sem > 0.8, radial <= max_radii, :603-605) + vote, all on the device
                from .producer import normalise_rgb
                if ev is None:
                    ev = FrameEvaluator(cls.cad_m * 1000, cls.keypoints_m[1:4, :] * 1000, sym, add_threshold[class_name] * 1000,
                                        frames_per_batch=frames_per_batch, image=(H, W), icp=icp, ctx=producer.ctx)
                images = normalise_rgb(np.stack([formats.read_rgb(cls.image_path(s)) for s in stems]))
                gt = np.stack([cls.pose_mm(s) for s in stems])
                res = ev.run_stage(producer, images, depth.view(np.int16) if depth.dtype == np.uint16 else depth, linemod_K, gt, cls.max_radii_dm)
                st = res["status"]
                bad = (st & api.RCV_ST_EMPTY_MASK) != 0
                if bad.any():
                    i, k = np.argwhere(bad)[0]
                    raise ValueError("zero-size array to reduction operation minimum which has no identity (frame %s, keypoint %d: empty mask)" % (stems[i], k + 1))
                if st.any():
                    raise api.RcvError("voting failed with status %s" % np.unique(st))
                for k, v in res.items():
                    acc.setdefault(k, []).append(v)
                continue
            radius = np.empty((len(stems), 3, H, W), np.float32)
            sem = np.empty_like(radius) if using_ckpts else None
            override_idx, override_val = [], []
            for i, s in enumerate(stems):
                for k in range(1, 4):
                    if using_ckpts:
                        sm, rd = producer(class_name, k, cls.image_path(s))
                        sem[i, k - 1], radius[i, k - 1] = sm, rd
                        continue
                    rd = cls.radial_est(s, k)
                    if rd.dtype == np.float32:
                        radius[i, k - 1] = rd
                        continue
                    # A map that is not float32 keeps its dtype through the reference's radius arithmetic (r*100/5), which
                    # the fused float32 path would not reproduce: vote this (frame, keypoint) through the exact drop-in surface
                    # (Accumulator_3D with the radii as they are, :612-628) and hand the fused path a float32 map with the same mask.
                    rd = np.where(rd <= cls.max_radii_dm[k - 1], rd, 0)
                    dm = depth[i] * np.where(rd != 0, 1, 0)
                    xyz_mm = shim.rgbd_to_point_cloud(linemod_K, dm)
                    override_idx.append((i, k - 1))
                    override_val.append(shim.Accumulator_3D(xyz_mm / 1000, rd[dm.nonzero()])[0])
                    radius[i, k - 1] = np.where(rd != 0, np.float32(1e-3), np.float32(0))
            if ev is None:
                ev = FrameEvaluator(cls.cad_m * 1000, cls.keypoints_m[1:4, :] * 1000, sym, add_threshold[class_name] * 1000, device=device,
                                    frames_per_batch=frames_per_batch, image=(H, W), icp=icp, max_grid=max_grid)
            gt = np.stack([cls.pose_mm(s) for s in stems])
            ov = None
            if override_idx:
                ii = np.array(override_idx)
                ov = ((torch.as_tensor(ii[:, 0]), torch.as_tensor(ii[:, 1])), np.array(override_val))
            res = ev.run(depth, radius, linemod_K, gt, max_radii=cls.max_radii_dm, sem=sem,
                         mask_flags=api.MASK_LM_CKPT if using_ckpts else api.MASK_LM_NPY, centres_override=ov)
            st = res["status"].copy()
            if override_idx:
                st[ii[:, 0], ii[:, 1]] = 0
            if ((st & api.RCV_ST_D_EXCEEDS_CAP) != 0).any():     # grids beyond max_grid: exact surface for those items, then the chain again
                rule = (lambda r, sm, k: (sm > 0.8) & (r <= cls.max_radii_dm[k])) if using_ckpts else (lambda r, sm, k: (r <= cls.max_radii_dm[k]) & (r != 0))
                _revote_oversized(res, (st & api.RCV_ST_D_EXCEEDS_CAP) != 0, radius, sem, depth, linemod_K, rule, shim, override_idx, override_val)
                ii = np.array(override_idx)
                ov = ((torch.as_tensor(ii[:, 0]), torch.as_tensor(ii[:, 1])), np.array(override_val))
                res = ev.run(depth, radius, linemod_K, gt, max_radii=cls.max_radii_dm, sem=sem,
                             mask_flags=api.MASK_LM_CKPT if using_ckpts else api.MASK_LM_NPY, centres_override=ov)
                st = res["status"].copy()
                st[ii[:, 0], ii[:, 1]] = 0
            bad = (st & api.RCV_ST_EMPTY_MASK) != 0
            if bad.any():   # the reference dies in Accumulator_3D on an empty cloud (ValueError from .min())
                i, k = np.argwhere(bad)[0]
                raise ValueError("zero-size array to reduction operation minimum which has no identity (frame %s, keypoint %d: empty mask)" % (stems[i], k + 1))
            if st.any():
                raise api.RcvError("voting failed with status %s" % np.unique(st))
            for k, v in res.items():
                acc.setdefault(k, []).append(v)
            if verbose:
                n = sum(len(a) for a in acc["passed_before"])
                print("Current ADD\\(s\\) of " + class_name + " before ICP: ", np.concatenate(acc["passed_before"]).sum() / n)
                if icp:
                    print("Currnet ADD\\(s\\) of " + class_name + " after ICP: ", np.concatenate(acc["passed_after"]).sum() / n)
    except Exception as e:   # noqa: BLE001 -- re-raised by sync_failure on every rank
        failure = e
    sync_failure(failure)
    out = {k: np.concatenate(v) for k, v in acc.items()}
    nb, na = reduce_counts([out["passed_before"].sum() if acc else 0, out["passed_after"].sum() if (acc and icp) else 0])
    n = len(cls.stems)
    out.update(frames=list(my_stems), n=n, add_before=float(nb / n) if n else float("nan"),
               add_after=float(na / n) if (n and icp) else float("nan"))
    if verbose:
        tag = "ADDs" if sym else "ADD"
        print(tag + " of " + class_name + " before ICP: ", out["add_before"])
        print(tag + " of " + class_name + " after ICP: ", out["add_after"])
    return out


def estimate_6d_pose_lm(opts):
    """Drop-in for AccumulatorSpace.estimate_6d_pose_lm(opts) (:495-744): opts.root_dataset, opts.using_ckpts [, opts.producer,
    opts.classes, opts.device, opts.frames_per_batch].  Prints the reference's summary lines and, unlike the reference (which
    returns None), returns {class_name: result dict}."""
    results = {}
    for class_name in getattr(opts, "classes", None) or lm_cls_names:
        print("Evaluation on ", class_name)
        results[class_name] = evaluate_lm_class(opts.root_dataset, class_name, using_ckpts=bool(getattr(opts, "using_ckpts", False)),
                                                producer=getattr(opts, "producer", None), device=_default_device(opts),
                                                frames_per_batch=getattr(opts, "frames_per_batch", 64), max_grid=getattr(opts, "max_grid", 256))
    return results


# ------------------------------------------------------------------------------------------------
# Occlusion LINEMOD (AccumulatorSpace.py:741-983)
# ------------------------------------------------------------------------------------------------
class LmoClass:
    """One class of the reference's Occlusion-LINEMOD layout (AccumulatorSpace.py:746, :769-785, :809-850):
        <root>/LINEMOD/<cls>/{<cls>.ply, Outside9.npy}
        <root>/OCCLUSION_LINEMOD/RGB-D/rgb_noseg/color_<NNNNN>.png, RGB-D/depth_noseg/depth_<NNNNN>.png   (depth in mm)
        <root>/OCCLUSION_LINEMOD/blender_poses/<cls>/pose<N>.npy
        <root>/OCCLUSION_LINEMOD/estRadialMap/<cls>/Out_pt<k>_dm/_<NNNNN>.npy
    `entries` is every directory entry of rgb_noseg (the reference's counter advances for each one, :962); `stems` are the
    frames that are evaluated: .png images whose pose and (npy branch) three radius maps exist (:809-817)."""

    def __init__(self, root_dataset, class_name, using_ckpts=False):
        self.name = class_name
        self.root = root_dataset + "OCCLUSION_LINEMOD/"
        pv = root_dataset + "LINEMOD/" + class_name + "/"
        self.cad_m = formats.read_ply_points(pv + class_name + ".ply")
        self.keypoints_m = formats.load_keypoints(pv + "Outside9.npy")
        self.max_radii_dm = max_radii_dm(self.cad_m, self.keypoints_m)
        self.entries = sorted(os.listdir(self.root + "RGB-D/rgb_noseg/"))
        self.stems = []
        for f in self.entries:
            if not f.endswith(".png"):
                continue
            stem = os.path.splitext(f)[0]
            ok = os.path.isfile(self.pose_path(stem))
            if not using_ckpts:
                ok = ok and all(os.path.isfile(self.radial_path(stem, k)) for k in (1, 2, 3))
            if ok:
                self.stems.append(stem)

    @staticmethod
    def index(stem):
        return int(stem[6:])                       # "color_00076" -> 76

    def image_path(self, stem):
        return self.root + "RGB-D/rgb_noseg/" + stem + ".png"

    def pose_path(self, stem):
        return self.root + "blender_poses/" + self.name + "/pose" + str(self.index(stem)) + ".npy"

    def radial_path(self, stem, k):
        return os.path.join(self.root, "estRadialMap", self.name, "Out_pt" + str(k) + "_dm", "_" + str(self.index(stem)).zfill(5) + ".npy")

    def depth(self, stem):
        return np.array(formats.read_depth(self.root + "RGB-D/depth_noseg/depth_" + stem[6:].zfill(5) + ".png"), dtype=np.float64)

    def pose_mm(self, stem):
        rt = formats.load_pose(self.pose_path(stem)).copy()
        rt[:, 3] = rt[:, 3] * 1000
        return rt


def evaluate_lmo_class(root_dataset, class_name, using_ckpts=False, producer=None, device=0, frames_per_batch=64, icp=True, verbose=True,
                       max_grid=256):
    """One class of estimate_6d_pose_lmo (AccumulatorSpace.py:741-983).  Differences from the LINEMOD evaluator, all the
    reference's: float64 depth images in mm (:833), mask rule `radial > 0` (npy, :849-851) or `sem >= 0.5` (ckpt, :837-840), a
    keypoint whose thresholded radius map is all zero is skipped and its row stays 0 (:858, :795), ICP runs with
    relative_fitness = relative_rmse = add_threshold * 1000 (:939-941), and the ADD(-S) ratios are taken over EVERY entry of the
    image directory, evaluated or not (:962)."""
    from . import AccumulatorSpace as shim
    cls = LmoClass(root_dataset, class_name, using_ckpts)
    sym = class_name in lm_syms
    thr = add_threshold[class_name] * 1000
    if using_ckpts and producer is None:
        raise ValueError("using_ckpts needs a producer(class_name, k, image_path) -> (sem, radial): the radius-map network is not part of rcvpose_b200")
    ev, acc = None, {}
    my_stems = shard_frames(cls.stems)
    verbose = verbose and _world()[0] == 0
    failure = None
    try:
        for b0 in range(0, len(my_stems), frames_per_batch):
            stems = my_stems[b0:b0 + frames_per_batch]
            depth = np.stack([cls.depth(s) for s in stems])
            H, W = depth.shape[1:]
            radius = np.empty((len(stems), 3, H, W), np.float32)
            sem = np.empty_like(radius) if using_ckpts else None
            zero_map = np.zeros((len(stems), 3), bool)
            override_idx, override_val = [], []
            for i, s in enumerate(stems):
                for k in range(1, 4):
                    if using_ckpts:
                        sm, rd = producer(class_name, k, cls.image_path(s))
                        sem[i, k - 1], radius[i, k - 1] = sm, rd
                        m = (np.asarray(sm) >= 0.5) & (np.asarray(rd) <= cls.max_radii_dm[k - 1])
                        zero_map[i, k - 1] = (np.asarray(rd) * m).max() == 0
                        continue
                    rd = np.load(cls.radial_path(s, k))
                    rd = np.where(rd <= cls.max_radii_dm[k - 1], rd, 0)
                    zero_map[i, k - 1] = rd.max() == 0
                    if rd.dtype == np.float32:
                        radius[i, k - 1] = rd
                        continue
                    # not float32: exact drop-in surface for this (frame, keypoint), as in evaluate_lm_class
                    radius[i, k - 1] = np.where(rd > 0, np.float32(1e-3), np.float32(0))
                    if not zero_map[i, k - 1]:
                        dm = depth[i] * np.where(rd > 0, 1, 0)
                        xyz_mm = shim.rgbd_to_point_cloud(linemod_K, dm)
                        override_idx.append((i, k - 1))
                        override_val.append(shim.Accumulator_3D(xyz_mm / 1000, rd[dm.nonzero()])[0])
            if ev is None:
                ev = FrameEvaluator(cls.cad_m * 1000, cls.keypoints_m[1:4, :] * 1000, sym, thr, device=device, frames_per_batch=frames_per_batch,
                                    image=(H, W), icp=icp, max_grid=max_grid)
            gt = np.stack([cls.pose_mm(s) for s in stems])
            ov = None
            if override_idx:
                ii = np.array(override_idx)
                ov = ((torch.as_tensor(ii[:, 0]), torch.as_tensor(ii[:, 1])), np.array(override_val))
            res = ev.run(depth, radius, linemod_K, gt, max_radii=cls.max_radii_dm, sem=sem, sem_threshold=0.5,
                         mask_flags=api.MASK_LMO_CKPT if using_ckpts else api.MASK_LMO_NPY, centres_override=ov, icp_rel_fitness=thr,
                         icp_rel_rmse=thr, zero_empty_centres=True)
            st = res["status"].copy()
            if override_idx:
                st[ii[:, 0], ii[:, 1]] = 0
            if ((st & api.RCV_ST_D_EXCEEDS_CAP) != 0).any():
                rule = (lambda r, sm, k: (sm >= 0.5) & (r <= cls.max_radii_dm[k])) if using_ckpts else (lambda r, sm, k: (r <= cls.max_radii_dm[k]) & (r > 0))
                _revote_oversized(res, (st & api.RCV_ST_D_EXCEEDS_CAP) != 0, radius, sem, depth, linemod_K, rule, shim, override_idx, override_val)
                ii = np.array(override_idx)
                ov = ((torch.as_tensor(ii[:, 0]), torch.as_tensor(ii[:, 1])), np.array(override_val))
                res = ev.run(depth, radius, linemod_K, gt, max_radii=cls.max_radii_dm, sem=sem, sem_threshold=0.5,
                             mask_flags=api.MASK_LMO_CKPT if using_ckpts else api.MASK_LMO_NPY, centres_override=ov, icp_rel_fitness=thr,
                             icp_rel_rmse=thr, zero_empty_centres=True)
                st = res["status"].copy()
                st[ii[:, 0], ii[:, 1]] = 0
            empty = (st & api.RCV_ST_EMPTY_MASK) != 0
            if (empty & ~zero_map).any():   # radii survive but no depth under them: the reference calls Accumulator_3D on an empty cloud
                i, k = np.argwhere(empty & ~zero_map)[0]
                raise ValueError("zero-size array to reduction operation minimum which has no identity (frame %s, keypoint %d: empty cloud)" % (stems[i], k + 1))
            if (st & ~api.RCV_ST_EMPTY_MASK).any():
                raise api.RcvError("voting failed with status %s" % np.unique(st))
            for k, v in res.items():
                acc.setdefault(k, []).append(v)
    except Exception as e:   # noqa: BLE001 -- re-raised by sync_failure on every rank
        failure = e
    sync_failure(failure)
    n = len(cls.entries)
    out = {k: np.concatenate(v) for k, v in acc.items()}
    nb, na = reduce_counts([out["passed_before"].sum() if acc else 0, out["passed_after"].sum() if (acc and icp) else 0])
    out.update(frames=list(my_stems), n=n, evaluated=len(cls.stems), add_before=nb / n if n else float("nan"),
               add_after=(na / n if icp else float("nan")) if n else float("nan"))
    if verbose:
        tag = "ADDs" if sym else "ADD"
        print(tag + " of " + class_name + " before ICP: ", out["add_before"])
        print(tag + " of " + class_name + " after ICP: ", out["add_after"])
    return out


def estimate_6d_pose_lmo(opts):
    """Drop-in for AccumulatorSpace.estimate_6d_pose_lmo(opts) (:741-983); returns {class_name: result dict}."""
    results = {}
    for class_name in getattr(opts, "classes", None) or lmo_cls_names:
        print(class_name)
        results[class_name] = evaluate_lmo_class(opts.root_dataset, class_name, using_ckpts=bool(getattr(opts, "using_ckpts", False)),
                                                 producer=getattr(opts, "producer", None), device=_default_device(opts),
                                                 frames_per_batch=getattr(opts, "frames_per_batch", 64), max_grid=getattr(opts, "max_grid", 256))
    return results


# ------------------------------------------------------------------------------------------------
# YCB-Video (AccumulatorSpace.py:976-1197)
# ------------------------------------------------------------------------------------------------
ycb_cls_names = {1: '002_master_chef_can', 2: '003_cracker_box', 3: '004_sugar_box', 4: '005_tomato_soup_can', 5: '006_mustard_bottle',
                 6: '007_tuna_fish_can', 7: '008_pudding_box', 8: '009_gelatin_box', 9: '010_potted_meat_can', 10: '011_banana',
                 11: '019_pitcher_base', 12: '021_bleach_cleanser', 13: '024_bowl', 14: '025_mug', 15: '035_power_drill',
                 16: '036_wood_block', 17: '037_scissors', 18: '040_large_marker', 19: '051_large_clamp', 20: '052_extra_large_clamp',
                 21: '061_foam_brick'}                                   # :21-41
ycb_syms = ['024_bowl', '036_wood_block', '051_large_clamp', '052_extra_large_clamp', '061_foam_brick']   # :44
ycb_auc_thresholds_m = [0, 0.02, 0.04, 0.06, 0.08, 0.1]                  # :978


def obb_diagonal(points):
    """Diagonal of the oriented bounding box open3d 0.14 builds from a point set (`get_oriented_bounding_box().extent`, :1118-1119):
    axes = eigenvectors of the covariance of the points, extent = the points' range along each axis.  (Restated from open3d's
    published algorithm -- open3d is not available here, so this figure is not pinned against it.)"""
    p = np.asarray(points, dtype=np.float64)
    c = p - p.mean(0)
    _, vec = np.linalg.eigh(c.T @ c / len(p))
    q = c @ vec
    ext = q.max(0) - q.min(0)
    return float(np.sqrt((ext ** 2).sum()))


def trapezoid_auc(x, y):
    """sklearn.metrics.auc(x, y) for increasing x (:1194-1195): the trapezoidal rule."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    return float(((x[1:] - x[:-1]) * (y[1:] + y[:-1]) / 2).sum())


class YcbClass:
    """One class of the layout estimate_6d_pose_ycb reads (:981-990, :1012-1016, :1044-1050):
        <root>/Split/<cls>/val.txt                      frame names "<cycle>_<idx>"
        <root>/models/<cls>/{points.xyz, Outside9.npy}  CAD points and keypoints, metres
        <root>/<cls>.hdf5, group JPEGImages             the class's frames; only the key names are used by the reference (:1011-1012)
        <root>/data/<cycle>/<idx>.mat                   poses (3,4,M) metres, cls_indexes, factor_depth, intrinsic_matrix (:1015)
        <root>/data/<cycle>/<idx>-color.png, <idx>-depth.png
    The frame list is `val.txt` intersected with the HDF5 keys when h5py is importable, otherwise with the frames whose files
    exist under data/ (h5py is not installed in this image); both in sorted order, the order h5py iterates keys in.
    `<idx>-meta.mat`, the name the YCB-Video distribution uses, is accepted as well."""

    def __init__(self, root_dataset, class_id):
        self.id, self.name = int(class_id), ycb_cls_names[int(class_id)]
        self.root = root_dataset
        self.cad_m = formats.read_xyz_points(root_dataset + "models/" + self.name + "/points.xyz")
        self.keypoints_m = formats.load_keypoints(root_dataset + "models/" + self.name + "/Outside9.npy")
        test = formats.read_split(root_dataset + "Split/" + self.name + "/val.txt")
        keys = None
        h5 = root_dataset + "/" + self.name + ".hdf5"
        if os.path.isfile(h5):
            try:
                import h5py
                with h5py.File(h5, "r") as f:
                    keys = set(f["JPEGImages/"].keys())
            except ImportError:
                keys = None
        self.stems = sorted(s for s in set(test) if s and (keys is None or s in keys) and os.path.isfile(self.meta_path(s)))

    def _base(self, stem):
        cycle, idx = stem.split("_")
        return self.root + "data/" + cycle + "/" + idx

    def meta_path(self, stem):
        b = self._base(stem)
        return b + ".mat" if os.path.isfile(b + ".mat") else b + "-meta.mat"

    def image_path(self, stem):
        return self._base(stem) + "-color.png"

    def meta(self, stem):
        return formats.load_ycb_meta(self.meta_path(stem))

    def depth_raw(self, stem):
        return np.asarray(formats.read_depth(self._base(stem) + "-depth.png"))


def evaluate_ycb_class(root_dataset, class_id, producer, device=0, frames_per_batch=32, icp=True, verbose=True, horn_keypoints=(1, 4),
                       max_grid=384):
    """One class of estimate_6d_pose_ycb (AccumulatorSpace.py:976-1197) as the reference INTENDS it.  The function cannot run as
    written; what is repaired here, and how:
      * `keypoint_count` is read before it is assigned (:1003) and indexes a 3-element list with 1..3 (:1044) and a (3,3) array
        with 1..3 (:1094): network k, keypoint k and row k - 1 belong together, as in the LINEMOD evaluator (:520-530, :654);
      * `RTGT = poses[:, :, np.where(cls_indexes == class_id)]` (:1016) indexes with a tuple of arrays: the pose of the object
        whose cls_index is class_id is meant (formats.pose_of);
      * Horn is given `keypoints[0:3]` (:1115) against estimates of keypoints 1..3: `horn_keypoints` selects the model rows,
        default (1, 4) = the keypoints that were voted for (pass (0, 3) for the literal rows);
      * the summary lines sit outside the class loop (:1194-1197) and the AUC counters are never reset: here they are per class.
    Kept as the reference has them: mask = sem > 0.8 only (:1049), depth / factor_depth in metres straight into
    rgbd_to_point_cloud with the frame's own intrinsic matrix (:1050-1057), Accumulator_3D on metres and decimetres (:1067),
    threshold = 1 % of the OBB diagonal (:1118-1119, :1141), ADD-S (minimum distance) for ycb_syms, ICP with
    max_correspondence_distance = the MEAN distance for every class and max_iteration = 2000000 (:1151-1158), AUC over
    [0, 0.1] m in steps of 0.02 divided by 0.1 (:1143-1150, :1194-1195).  Surviving pixel = sem > 0.8 and depth != 0 for both lists
    (the reference's two lists mis-align when the mask covers a depth hole).
    `producer(class_name, k, image_path) -> (sem, radial)` supplies the maps of keypoint k = 1..3 (decimetres)."""
    cls = YcbClass(root_dataset, class_id)
    sym = cls.name in ycb_syms
    if producer is None:
        raise ValueError("the YCB evaluator has only a checkpoint branch: it needs a producer(class_name, k, image_path) -> (sem, radial)")
    thr_mm = obb_diagonal(cls.cad_m) * 0.01 * 1000
    k0, k1 = horn_keypoints
    ev, acc = None, {}
    my_stems = shard_frames(cls.stems)
    verbose = verbose and _world()[0] == 0
    failure = None
    try:
        for b0 in range(0, len(my_stems), frames_per_batch):
            stems = my_stems[b0:b0 + frames_per_batch]
            metas = [cls.meta(s) for s in stems]
            factor = metas[0]["factor_depth"]
            if any(m["factor_depth"] != factor for m in metas):
                raise ValueError("frames of one batch must share factor_depth (got %s)" % sorted({m["factor_depth"] for m in metas}))
            depth = np.stack([cls.depth_raw(s) for s in stems])
            H, W = depth.shape[1:]
            radius = np.empty((len(stems), 3, H, W), np.float32)
            sem = np.empty_like(radius)
            gt = np.zeros((len(stems), 3, 4))
            for i, s in enumerate(stems):
                rt = formats.pose_of(metas[i], cls.id)
                if rt is None:
                    raise ValueError("frame %s does not contain object %d (%s)" % (s, cls.id, cls.name))
                gt[i] = rt
                gt[i, :, 3] *= 1000
                for k in range(1, 4):
                    sem[i, k - 1], radius[i, k - 1] = producer(cls.name, k, cls.image_path(s))
            if ev is None:
                ev = FrameEvaluator(cls.cad_m * 1000, cls.keypoints_m[k0:k1, :] * 1000, sym, thr_mm, device=device, frames_per_batch=frames_per_batch,
                                    image=(H, W), icp=icp, icp_max_iter=2000000, max_grid=max_grid)
            Ks = np.stack([m["intrinsic_matrix"] for m in metas])
            d = depth.view(np.int16) if depth.dtype == np.uint16 else depth.astype(np.float64)
            res = ev.run(d, radius, Ks, gt, sem=sem, mask_flags=api.MASK_YCB, sem_threshold=0.8, depth_div=factor, xyz_div=1.0, scene_scale=1000.0,
                         icp_threshold_mean=True)
            st = res["status"]
            if ((st & api.RCV_ST_EMPTY_MASK) != 0).any():
                i, k = np.argwhere((st & api.RCV_ST_EMPTY_MASK) != 0)[0]
                raise ValueError("zero-size array to reduction operation minimum which has no identity (frame %s, keypoint %d: empty mask)" % (stems[i], k + 1))
            if st.any():
                raise api.RcvError("voting failed with status %s" % np.unique(st))
            for k, v in res.items():
                acc.setdefault(k, []).append(v)
    except Exception as e:   # noqa: BLE001 -- re-raised by sync_failure on every rank
        failure = e
    sync_failure(failure)
    out = {k: np.concatenate(v) for k, v in acc.items()}
    n = len(cls.stems)
    thr = np.array(ycb_auc_thresholds_m) * 1000
    counts = [out["passed_before"].sum() if acc else 0, out["passed_after"].sum() if (acc and icp) else 0]
    counts += [(out["dist_before"] <= t).sum() if acc else 0 for t in thr]
    counts += [(out["dist_after"] <= t).sum() if (acc and icp) else 0 for t in thr]
    counts = reduce_counts(counts)
    nan = float("nan")
    out.update(frames=list(my_stems), n=n, threshold_mm=thr_mm, add_before=counts[0] / n if n else nan, add_after=(counts[1] / n if icp else nan) if n else nan,
               auc_before=trapezoid_auc(ycb_auc_thresholds_m, np.array(counts[2:8]) / n) / 0.1 if n else nan,
               auc_after=(trapezoid_auc(ycb_auc_thresholds_m, np.array(counts[8:14]) / n) / 0.1 if icp else nan) if n else nan)
    if verbose:
        print('ADD\\(s\\) AUC of ' + cls.name + ' before ICP: ', out["auc_before"])
        print('ADD\\(s\\) AUC of ' + cls.name + ' after ICP: ', out["auc_after"])
        print('ADD\\(s\\) of ' + cls.name + ' before ICP: ', out["add_before"])
        print('ADD\\(s\\) of ' + cls.name + ' after ICP: ', out["add_after"])
    return out


def estimate_6d_pose_ycb(opts):
    """Drop-in for AccumulatorSpace.estimate_6d_pose_ycb(opts) (:976-1197), repaired as evaluate_ycb_class documents:
    opts.root_dataset, opts.producer [, opts.classes (ids), opts.device, opts.frames_per_batch, opts.horn_keypoints].
    Returns {class_name: result dict}."""
    results = {}
    for class_id in getattr(opts, "classes", None) or list(ycb_cls_names):
        print(ycb_cls_names[class_id])
        results[ycb_cls_names[class_id]] = evaluate_ycb_class(opts.root_dataset, class_id, getattr(opts, "producer", None), device=_default_device(opts),
                                                              frames_per_batch=getattr(opts, "frames_per_batch", 32),
                                                              horn_keypoints=getattr(opts, "horn_keypoints", (1, 4)))
    return results
