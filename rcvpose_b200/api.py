"""Host-side API over librcvvote.so: PyTorch owns device memory and streams, CUDA does the work.

`VoteContext` is a thin object wrapper of the C ABI (include/rcvvote.h); every method takes CUDA
tensors (or, for the `_host` methods, NumPy arrays) and launches asynchronously on torch's
current stream.  Nothing here computes votes, peaks or poses on the CPU.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (RCV_F32, RCV_F64, RCV_U16, RCV_POLICY_LM, RCV_POLICY_YCBGEN, RCV_MASK_RADIUS_NONZERO,  # noqa: F401
                   RCV_MASK_RADIUS_POSITIVE, RCV_MASK_SEM_GT, RCV_MASK_SEM_GE, RCV_MASK_MAX_RADIUS, RCV_ST_OK,
                   RCV_ST_EMPTY_MASK, RCV_ST_BAD_GRID, RCV_ST_D_EXCEEDS_CAP, RCV_ST_POINT_OVERFLOW, RCV_ST_UNIT_OVERFLOW,
                   RCV_ST_VOLUME_SKIPPED)

_DEPTH_DTYPES = {torch.uint16: RCV_U16, torch.int16: RCV_U16, torch.float32: RCV_F32, torch.float64: RCV_F64}
_NP_DEPTH_DTYPES = {np.dtype(np.uint16): RCV_U16, np.dtype(np.float32): RCV_F32, np.dtype(np.float64): RCV_F64}

# mask rules of the reference's three evaluators (AccumulatorSpace.py)
MASK_LM_NPY = RCV_MASK_MAX_RADIUS | RCV_MASK_RADIUS_NONZERO       # :612-618
MASK_LM_CKPT = RCV_MASK_MAX_RADIUS | RCV_MASK_SEM_GT              # :603-610  (sem > 0.8)
MASK_LMO_NPY = RCV_MASK_MAX_RADIUS | RCV_MASK_RADIUS_POSITIVE     # :849-851
MASK_LMO_CKPT = RCV_MASK_MAX_RADIUS | RCV_MASK_SEM_GE             # :837-840  (sem >= 0.5)
MASK_YCB = RCV_MASK_SEM_GT                                        # :1049-1053 (sem > 0.8)


class RcvError(RuntimeError):
    pass


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_cuda(t, dtype=None, name="tensor"):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.is_contiguous()):
        raise RcvError("%s must be a contiguous CUDA tensor" % name)
    if dtype is not None and t.dtype != dtype:
        raise RcvError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))


class VoteContext:
    """One context per (device, stream); not re-entrant.  Capacities are fixed here (no allocation
    on the hot calls)."""

    def __init__(self, device=0, max_items=64, max_points_total=1 << 22, max_grid=256, max_units=0, image=None, max_model_points=0,
                 head_items=0):
        if not torch.cuda.is_available():
            raise RcvError("no CUDA device: rcvpose_b200 has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device if isinstance(device, int) else device.index or 0)
        # image = (H, W) of the frames entry points, max_model_points = CAD size of the ADD / ICP entry points, head_items = items of
        # the fused head: given here, their scratch is allocated now and never on a hot call
        cfg = _lib.rcv_config(_lib.RCV_ABI_VERSION, int(max_items), int(max_points_total), int(max_grid), int(max_units),
                              int(image[0]) * int(image[1]) if image else 0, int(max_model_points), int(head_items))
        h = C.c_void_p()
        rc = self.lib.rcv_create(self.device.index, C.byref(cfg), C.byref(h))
        if rc != 0:
            raise RcvError("rcv_create failed (%d): %s" % (rc, self.lib.rcv_last_error(None).decode()))
        self.h = h
        self.max_items, self.max_grid, self.max_points_total = int(max_items), int(max_grid), int(max_points_total)

    def close(self):
        if getattr(self, "h", None):
            self.lib.rcv_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise RcvError("librcvvote error %d: %s" % (rc, self.lib.rcv_last_error(self.h).decode()))

    @property
    def launches(self):
        return int(self.lib.rcv_launch_count(self.h))

    @property
    def last_h2d_bytes(self):
        """Bytes the last vote_frames_host call moved host -> device (after cropping the images to their non-zero depth rows)."""
        return int(self.lib.rcv_last_h2d_bytes(self.h))

    def last_vote_kernel_ms(self):
        return float(self.lib.rcv_last_vote_kernel_ms(self.h))

    def vote_kernel_times(self, n):
        """Device durations (ms) of the vote kernel in the last n (<=64) calls, oldest first (syncs)."""
        buf = (C.c_float * 64)()
        got = self.lib.rcv_vote_kernel_times(self.h, buf, int(n))
        if got < 0:
            self._ck(got)
        return [float(buf[i]) for i in range(got)]

    def measure_smem_atomic_peak(self):
        """Conflict-free shared-memory atomics per second on this GPU (built-in micro-benchmark)."""
        v = C.c_double()
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_ubench_smem_atomics(self.h, C.byref(v)))
        return float(v.value)

    # ---- rgbd_to_point_cloud (AccumulatorSpace.py:77-85) ----
    def backproject(self, K, depth):
        _check_cuda(K, torch.float64, "K")
        _check_cuda(depth, None, "depth")
        H, W = depth.shape
        xyz = torch.empty((H * W, 3), dtype=torch.float64, device=depth.device)
        n = torch.zeros(1, dtype=torch.int32, device=depth.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_backproject(self.h, _ptr(K), _ptr(depth), _DEPTH_DTYPES[depth.dtype], H, W, _ptr(xyz), H * W, _ptr(n),
                                              _stream()))
        return xyz[: int(n.item())]

    # ---- Accumulator_3D (AccumulatorSpace.py:373-419), batched ----
    def vote_points(self, xyz, radii, item_offsets=None, acc_unit=5.0, radius_scale=100.0, policy=RCV_POLICY_LM, want_volume=False,
                    volume_capacity=None):
        _check_cuda(xyz, torch.float64, "xyz")
        _check_cuda(radii, None, "radii")
        if radii.dtype not in (torch.float32, torch.float64):
            raise RcvError("radii must be float32 or float64")
        dev = xyz.device
        if item_offsets is None:
            item_offsets = torch.tensor([0, xyz.shape[0]], dtype=torch.int64, device=dev)
        _check_cuda(item_offsets, torch.int64, "item_offsets")
        B = item_offsets.numel() - 1
        out = dict(centre_mm=torch.empty((B, 3), dtype=torch.float64, device=dev), peak=torch.empty(B, dtype=torch.int32, device=dev),
                   votes=torch.empty(B, dtype=torch.int64, device=dev), grid=torch.empty(B, dtype=torch.int32, device=dev),
                   zero_boundary=torch.empty(B, dtype=torch.int32, device=dev), status=torch.empty(B, dtype=torch.int32, device=dev))
        vol = None
        if want_volume:
            cap = int(volume_capacity if volume_capacity is not None else self.max_grid ** 3)
            vol = torch.zeros(cap, dtype=torch.int32, device=dev)
        vp = _lib.rcv_vote_params(float(acc_unit), float(radius_scale), int(policy), RCV_F32 if radii.dtype == torch.float32 else RCV_F64)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_vote_points(self.h, _ptr(xyz), _ptr(radii), _ptr(item_offsets), B, C.byref(vp), _ptr(out["centre_mm"]),
                                              _ptr(out["peak"]), _ptr(out["votes"]), _ptr(out["grid"]), _ptr(out["zero_boundary"]),
                                              _ptr(out["status"]), _ptr(vol), vol.numel() if vol is not None else 0, _stream()))
        if vol is not None:
            out["volume_flat"] = vol
        return out

    # ---- fused mask + back-projection + Accumulator_3D over frames x keypoints ----
    def vote_frames(self, depth, radius, K, sem=None, max_radii=None, mask_flags=MASK_LM_NPY, sem_threshold=0.8, depth_div=1.0,
                    xyz_div=1000.0, acc_unit=5.0, radius_scale=100.0, policy=RCV_POLICY_LM):
        _check_cuda(depth, None, "depth")
        _check_cuda(radius, torch.float32, "radius")
        _check_cuda(K, torch.float64, "K")
        B, Kp, H, W = radius.shape
        if tuple(depth.shape) != (B, H, W):
            raise RcvError("depth must be (B,H,W) matching radius (B,Kp,H,W)")
        if sem is not None:
            _check_cuda(sem, torch.float32, "sem")
        if max_radii is not None:
            _check_cuda(max_radii, torch.float64, "max_radii")
        dev = radius.device
        fp = _lib.rcv_frame_params(H, W, _DEPTH_DTYPES[depth.dtype], float(depth_div), float(xyz_div), int(mask_flags), float(sem_threshold),
                                   9 if (K.dim() == 3 and K.shape[0] == B) else 0,
                                   Kp if (max_radii is not None and max_radii.dim() == 2 and max_radii.shape[0] == B) else 0)
        vp = _lib.rcv_vote_params(float(acc_unit), float(radius_scale), int(policy), RCV_F32)
        out = dict(centre_mm=torch.empty((B, Kp, 3), dtype=torch.float64, device=dev), peak=torch.empty((B, Kp), dtype=torch.int32, device=dev),
                   votes=torch.empty((B, Kp), dtype=torch.int64, device=dev), n_points=torch.empty((B, Kp), dtype=torch.int32, device=dev),
                   grid=torch.empty((B, Kp), dtype=torch.int32, device=dev), status=torch.empty((B, Kp), dtype=torch.int32, device=dev))
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_vote_frames(self.h, B, Kp, _ptr(depth), _ptr(radius), _ptr(sem), _ptr(K), _ptr(max_radii), C.byref(fp),
                                              C.byref(vp), _ptr(out["centre_mm"]), _ptr(out["peak"]), _ptr(out["votes"]), _ptr(out["n_points"]),
                                              _ptr(out["grid"]), _ptr(out["status"]), _stream()))
        return out

    def vote_frames_host(self, depth, radius, K, sem=None, max_radii=None, mask_flags=MASK_LM_NPY, sem_threshold=0.8, depth_div=1.0,
                         xyz_div=1000.0, acc_unit=5.0, radius_scale=100.0, policy=RCV_POLICY_LM, frames_per_chunk=256, out=None):
        """HOST arrays in, HOST arrays out (NumPy or pinned CPU tensors' .numpy()); H2D/D2H inside."""
        B, Kp, H, W = radius.shape
        assert radius.dtype == np.float32 and radius.flags.c_contiguous and depth.flags.c_contiguous and tuple(depth.shape) == (B, H, W)
        K = np.ascontiguousarray(K, dtype=np.float64)
        mr = np.ascontiguousarray(max_radii, dtype=np.float64) if max_radii is not None else None
        fp = _lib.rcv_frame_params(H, W, _NP_DEPTH_DTYPES[depth.dtype], float(depth_div), float(xyz_div), int(mask_flags), float(sem_threshold),
                                   9 if (K.ndim == 3 and K.shape[0] == B) else 0,
                                   Kp if (mr is not None and mr.ndim == 2 and mr.shape[0] == B) else 0)
        vp = _lib.rcv_vote_params(float(acc_unit), float(radius_scale), int(policy), RCV_F32)
        if out is None:
            out = dict(centre_mm=np.empty((B, Kp, 3), np.float64), peak=np.empty((B, Kp), np.int32), votes=np.empty((B, Kp), np.int64),
                       n_points=np.empty((B, Kp), np.int32), grid=np.empty((B, Kp), np.int32), status=np.empty((B, Kp), np.int32))
        p = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None  # noqa: E731
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_vote_frames_host(self.h, B, Kp, p(depth), p(radius), p(sem), p(K), p(mr), C.byref(fp), C.byref(vp),
                                                   p(out["centre_mm"]), p(out["peak"]), p(out["votes"]), p(out["n_points"]), p(out["grid"]),
                                                   p(out["status"]), int(frames_per_chunk), _stream()))
        return out

    # ---- argwhere(V == V.max())[0] (AccumulatorSpace.py:406) ----
    def argmax_volume(self, volume):
        _check_cuda(volume, torch.int32, "volume")
        D = volume.shape[0]
        idx = torch.empty(3, dtype=torch.int32, device=volume.device)
        mx = torch.empty(1, dtype=torch.int32, device=volume.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_argmax_volume(self.h, _ptr(volume), D, _ptr(idx), _ptr(mx), _stream()))
        return idx, mx

    # ---- HornPoseFitting.lmshorn, batched (util/horn.py:75-181) ----
    def horn_batch(self, model, est):
        """model (n,3) or (B,n,3), est (B,n,3) float64 CUDA -> RT (B,4,4)."""
        _check_cuda(model, torch.float64, "model")
        _check_cuda(est, torch.float64, "est")
        B, n, _ = est.shape
        stride = 3 * n if model.dim() == 3 else 0
        RT = torch.empty((B, 4, 4), dtype=torch.float64, device=est.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_horn_batch(self.h, _ptr(model), stride, _ptr(est), n, B, _ptr(RT), _stream()))
        return RT

    # ---- ADD(-S) distance before ICP (AccumulatorSpace.py:664-702) ----
    def add_metric(self, model_mm, RT_est, RT_gt):
        """model_mm (M,3), RT_est / RT_gt (B,4,4) float64 CUDA (translations in the unit of model_mm) -> (mean, min) (B,) float64:
        mean and minimum over the ground-truth-transformed CAD points of the distance to the nearest estimate-transformed
        CAD point.  The reference thresholds the mean, or the minimum for its symmetric classes (:688-695)."""
        _check_cuda(model_mm, torch.float64, "model_mm")
        _check_cuda(RT_est, torch.float64, "RT_est")
        _check_cuda(RT_gt, torch.float64, "RT_gt")
        B = RT_est.shape[0]
        if RT_est.shape[1:] != (4, 4) or RT_gt.shape != RT_est.shape or model_mm.dim() != 2 or model_mm.shape[1] != 3:
            raise RcvError("add_metric: model_mm (M,3), RT_est and RT_gt (B,4,4)")
        mean = torch.empty(B, dtype=torch.float64, device=RT_est.device)
        mn = torch.empty(B, dtype=torch.float64, device=RT_est.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_add_metric_batch(self.h, _ptr(model_mm), int(model_mm.shape[0]), _ptr(RT_est), _ptr(RT_gt), B, _ptr(mean),
                                                   _ptr(mn), _stream()))
        return mean, mn

    # ---- xyz_mm_icp: union of a frame's masked clouds (AccumulatorSpace.py:620-625) ----
    def scene_clouds(self, depth, radius, K, sem=None, max_radii=None, mask_flags=MASK_LM_NPY, sem_threshold=0.8, depth_div=1.0, scale=1.0,
                     capacity=None):
        """Same maps and mask rules as vote_frames -> (xyz (N,3) float64 CUDA, offsets (B+1,) int64 CUDA, status (B,) int32 CUDA):
        frame f's scene cloud is xyz[offsets[f]:offsets[f+1]] = rgbd_to_point_cloud(K, depth * (mask_1 | ... | mask_Kp)) * scale."""
        _check_cuda(depth, None, "depth")
        _check_cuda(radius, torch.float32, "radius")
        _check_cuda(K, torch.float64, "K")
        B, Kp, H, W = radius.shape
        if tuple(depth.shape) != (B, H, W):
            raise RcvError("depth must be (B,H,W) matching radius (B,Kp,H,W)")
        if sem is not None:
            _check_cuda(sem, torch.float32, "sem")
        if max_radii is not None:
            _check_cuda(max_radii, torch.float64, "max_radii")
        dev = radius.device
        cap = int(capacity if capacity is not None else min(self.max_points_total, B * H * W))
        fp = _lib.rcv_frame_params(H, W, _DEPTH_DTYPES[depth.dtype], float(depth_div), 1000.0, int(mask_flags), float(sem_threshold),
                                   9 if (K.dim() == 3 and K.shape[0] == B) else 0,
                                   Kp if (max_radii is not None and max_radii.dim() == 2 and max_radii.shape[0] == B) else 0)
        xyz = torch.empty((cap, 3), dtype=torch.float64, device=dev)
        offsets = torch.empty(B + 1, dtype=torch.int64, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_scene_clouds(self.h, B, Kp, _ptr(depth), _ptr(radius), _ptr(sem), _ptr(K), _ptr(max_radii), C.byref(fp),
                                               float(scale), _ptr(xyz), cap, _ptr(offsets), _ptr(status), _stream()))
        return xyz, offsets, status

    def scene_clouds_last(self, depth, K, n_kpts, depth_div=1.0, scale=1.0, capacity=None):
        """scene_clouds for the frames of the most recent vote_frames / head_vote_frames call on this context, from the survival
        bits that call left behind (no map is read again).  depth (B,H,W), K (3,3) or (B,3,3) as in that call."""
        _check_cuda(depth, None, "depth")
        _check_cuda(K, torch.float64, "K")
        B, H, W = depth.shape
        cap = int(capacity if capacity is not None else min(self.max_points_total, B * H * W))
        fp = _lib.rcv_frame_params(H, W, _DEPTH_DTYPES[depth.dtype], float(depth_div), 1000.0, 0, 0.0, 9 if (K.dim() == 3 and K.shape[0] == B) else 0, 0)
        xyz = torch.empty((cap, 3), dtype=torch.float64, device=depth.device)
        offsets = torch.empty(B + 1, dtype=torch.int64, device=depth.device)
        status = torch.empty(B, dtype=torch.int32, device=depth.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_scene_clouds_last(self.h, B, int(n_kpts), _ptr(depth), _ptr(K), C.byref(fp), float(scale), _ptr(xyz), cap,
                                                    _ptr(offsets), _ptr(status), _stream()))
        return xyz, offsets, status

    # ---- open3d registration_icp, point to point (AccumulatorSpace.py:704-718) ----
    def icp(self, model, scene, scene_offsets, RT_init, max_dist, max_iter=30, rel_fitness=1e-6, rel_rmse=1e-6):
        """model (M,3), scene (N,3), scene_offsets (B+1,) int64, RT_init (B,4,4), max_dist (B,) -- all CUDA, float64 unless noted ->
        dict(RT (B,4,4), fitness (B,), rmse (B,), iters (B,) int32): open3d's reg.transformation / fitness / inlier_rmse."""
        for t, n in ((model, "model"), (scene, "scene"), (RT_init, "RT_init"), (max_dist, "max_dist")):
            _check_cuda(t, torch.float64, n)
        _check_cuda(scene_offsets, torch.int64, "scene_offsets")
        B = RT_init.shape[0]
        if RT_init.shape[1:] != (4, 4) or max_dist.numel() != B or scene_offsets.numel() != B + 1 or model.dim() != 2 or model.shape[1] != 3:
            raise RcvError("icp: model (M,3), scene (N,3), scene_offsets (B+1,), RT_init (B,4,4), max_dist (B,)")
        dev = RT_init.device
        out = dict(RT=torch.empty((B, 4, 4), dtype=torch.float64, device=dev), fitness=torch.empty(B, dtype=torch.float64, device=dev),
                   rmse=torch.empty(B, dtype=torch.float64, device=dev), iters=torch.empty(B, dtype=torch.int32, device=dev))
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_icp_batch(self.h, _ptr(model), int(model.shape[0]), _ptr(scene), _ptr(scene_offsets), _ptr(RT_init),
                                            _ptr(max_dist), B, int(max_iter), float(rel_fitness), float(rel_rmse), _ptr(out["RT"]),
                                            _ptr(out["fitness"]), _ptr(out["rmse"]), _ptr(out["iters"]), _stream()))
        return out

    def head_1x1(self, up, weight, bias):
        """conv8 of the reference's producer (models/fcnresnet.py:118,187-189) on the tensor cores.
        up (B,32,H,W) bfloat16 NCHW, weight (2,32[,1,1]) and bias (2,) float32 -> out (B,2,H,W) float32:
        out[:,0] = seg map, out[:,1] = radius map."""
        if isinstance(up, torch.Tensor):
            up = up.contiguous()
        _check_cuda(up, torch.bfloat16, "up")
        B, Cin, H, W = up.shape
        if Cin != 32:
            raise ValueError("head_1x1: the head has 32 input channels")
        weight = weight.reshape(2, 32).to(device=up.device, dtype=torch.float32).contiguous()
        bias = bias.reshape(2).to(device=up.device, dtype=torch.float32).contiguous()
        out = torch.empty((B, 2, H, W), dtype=torch.float32, device=up.device)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_head_1x1(self.h, _ptr(up), _ptr(weight), _ptr(bias), _ptr(out), B, H * W, _stream()))
        return out

    # ---- conv8 -> mask rule -> voting, fused (models/fcnresnet.py:187-189 + AccumulatorSpace.py:603-656) ----
    def head_vote_frames(self, up, weight, bias, depth, K, max_radii=None, mask_flags=MASK_LM_CKPT, sem_threshold=0.8, depth_div=1.0,
                         xyz_div=1000.0, acc_unit=5.0, radius_scale=100.0, policy=RCV_POLICY_LM, want_radius=False):
        """up (B,Kp,32,H,W) bfloat16, weight (Kp,2,32[,1,1]), bias (Kp,2), depth (B,H,W), K (3,3)/(B,3,3) float64 -- CUDA.  Same outputs as
        vote_frames (bit-identical to head_1x1 + vote_frames with sem = the head's seg plane); out["radius"] (B,Kp,H,W) if want_radius."""
        if isinstance(up, torch.Tensor):
            up = up.contiguous()
        _check_cuda(up, torch.bfloat16, "up")
        _check_cuda(depth, None, "depth")
        _check_cuda(K, torch.float64, "K")
        B, Kp, Cin, H, W = up.shape
        if Cin != 32 or tuple(depth.shape) != (B, H, W):
            raise RcvError("head_vote_frames: up (B,Kp,32,H,W), depth (B,H,W)")
        dev = up.device
        weight = weight.reshape(Kp, 2, 32).to(device=dev, dtype=torch.float32).contiguous()
        bias = bias.reshape(Kp, 2).to(device=dev, dtype=torch.float32).contiguous()
        if max_radii is not None:
            _check_cuda(max_radii, torch.float64, "max_radii")
        fp = _lib.rcv_frame_params(H, W, _DEPTH_DTYPES[depth.dtype], float(depth_div), float(xyz_div), int(mask_flags), float(sem_threshold),
                                   9 if (K.dim() == 3 and K.shape[0] == B) else 0,
                                   Kp if (max_radii is not None and max_radii.dim() == 2 and max_radii.shape[0] == B) else 0)
        vp = _lib.rcv_vote_params(float(acc_unit), float(radius_scale), int(policy), RCV_F32)
        out = dict(centre_mm=torch.empty((B, Kp, 3), dtype=torch.float64, device=dev), peak=torch.empty((B, Kp), dtype=torch.int32, device=dev),
                   votes=torch.empty((B, Kp), dtype=torch.int64, device=dev), n_points=torch.empty((B, Kp), dtype=torch.int32, device=dev),
                   grid=torch.empty((B, Kp), dtype=torch.int32, device=dev), status=torch.empty((B, Kp), dtype=torch.int32, device=dev))
        rad = torch.empty((B, Kp, H, W), dtype=torch.float32, device=dev) if want_radius else None
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_head_vote_frames(self.h, B, Kp, _ptr(up), _ptr(weight), _ptr(bias), _ptr(depth), _ptr(K), _ptr(max_radii),
                                                   C.byref(fp), C.byref(vp), _ptr(out["centre_mm"]), _ptr(out["peak"]), _ptr(out["votes"]),
                                                   _ptr(out["n_points"]), _ptr(out["grid"]), _ptr(out["status"]), _ptr(rad), _stream()))
        if rad is not None:
            out["radius"] = rad
        return out

    # ---- conv7 + BN + ReLU + conv8 in one kernel (models/fcnresnet.py:114-118, :183-189) ----
    @staticmethod
    def _nhwc(x, name):
        """(B,64,H,W) bfloat16 CUDA tensor -> the same tensor in channels_last memory (no copy if it already is)."""
        if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] == 64):
            raise RcvError("%s: (B,64,H,W) bfloat16 CUDA tensor" % name)
        return x.contiguous(memory_format=torch.channels_last)

    def conv7_head(self, x, w7, bn_scale, bn_shift, w8, b8):
        """x (B,64,H,W) bfloat16 (the output of up1; channels_last memory is used as is), w7 (32,64,3,3), bn_scale / bn_shift (32,)
        (BatchNorm folded: producer.fold_conv7_bn), w8 (2,32[,1,1]), b8 (2,) -> out (B,2,H,W) float32: seg, radial."""
        x = self._nhwc(x, "x")
        B, _, H, W = x.shape
        dev = x.device
        f = lambda t, shape: t.detach().reshape(shape).to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        w7, bn_scale, bn_shift, w8, b8 = f(w7, (32, 64, 3, 3)), f(bn_scale, (32,)), f(bn_shift, (32,)), f(w8, (2, 32)), f(b8, (2,))
        out = torch.empty((B, 2, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_conv7_head(self.h, _ptr(x), _ptr(w7), _ptr(bn_scale), _ptr(bn_shift), _ptr(w8), _ptr(b8), _ptr(out), B, H, W, _stream()))
        return out

    def conv7_head_vote_frames(self, xs, w7, bn_scale, bn_shift, w8, b8, depth, K, max_radii=None, mask_flags=MASK_LM_CKPT, sem_threshold=0.8,
                               depth_div=1.0, xyz_div=1000.0, acc_unit=5.0, radius_scale=100.0, policy=RCV_POLICY_LM, want_radius=False):
        """xs: Kp tensors (B,64,H,W) bfloat16 (up1 output of each keypoint network), w7 (Kp,32,64,3,3), bn_scale / bn_shift (Kp,32),
        w8 (Kp,2,32[,1,1]), b8 (Kp,2), depth (B,H,W), K (3,3)/(B,3,3) float64 -- CUDA.  Outputs as head_vote_frames."""
        xs = [self._nhwc(x, "xs[%d]" % i) for i, x in enumerate(xs)]
        Kp = len(xs)
        B, _, H, W = xs[0].shape
        _check_cuda(depth, None, "depth")
        _check_cuda(K, torch.float64, "K")
        if any(tuple(x.shape) != (B, 64, H, W) for x in xs) or tuple(depth.shape) != (B, H, W):
            raise RcvError("conv7_head_vote_frames: xs Kp x (B,64,H,W), depth (B,H,W)")
        dev = xs[0].device
        f = lambda t, shape: t.detach().reshape(shape).to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        w7, bn_scale, bn_shift, w8, b8 = f(w7, (Kp, 32, 64, 3, 3)), f(bn_scale, (Kp, 32)), f(bn_shift, (Kp, 32)), f(w8, (Kp, 2, 32)), f(b8, (Kp, 2))
        if max_radii is not None:
            _check_cuda(max_radii, torch.float64, "max_radii")
        fp = _lib.rcv_frame_params(H, W, _DEPTH_DTYPES[depth.dtype], float(depth_div), float(xyz_div), int(mask_flags), float(sem_threshold),
                                   9 if (K.dim() == 3 and K.shape[0] == B) else 0,
                                   Kp if (max_radii is not None and max_radii.dim() == 2 and max_radii.shape[0] == B) else 0)
        vp = _lib.rcv_vote_params(float(acc_unit), float(radius_scale), int(policy), RCV_F32)
        out = dict(centre_mm=torch.empty((B, Kp, 3), dtype=torch.float64, device=dev), peak=torch.empty((B, Kp), dtype=torch.int32, device=dev),
                   votes=torch.empty((B, Kp), dtype=torch.int64, device=dev), n_points=torch.empty((B, Kp), dtype=torch.int32, device=dev),
                   grid=torch.empty((B, Kp), dtype=torch.int32, device=dev), status=torch.empty((B, Kp), dtype=torch.int32, device=dev))
        rad = torch.empty((B, Kp, H, W), dtype=torch.float32, device=dev) if want_radius else None
        xp = (C.c_void_p * Kp)(*[x.data_ptr() for x in xs])
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_conv7_head_vote_frames(self.h, B, Kp, xp, _ptr(w7), _ptr(bn_scale), _ptr(bn_shift), _ptr(w8), _ptr(b8), _ptr(depth),
                                                         _ptr(K), _ptr(max_radii), C.byref(fp), C.byref(vp), _ptr(out["centre_mm"]), _ptr(out["peak"]),
                                                         _ptr(out["votes"]), _ptr(out["n_points"]), _ptr(out["grid"]), _ptr(out["status"]), _ptr(rad),
                                                         _stream()))
        if rad is not None:
            out["radius"] = rad
        return out

    def horn_batch_host(self, model, est):
        model = np.ascontiguousarray(model, dtype=np.float64)
        est = np.ascontiguousarray(est, dtype=np.float64)
        B, n, _ = est.shape
        RT = np.empty((B, 4, 4), np.float64)
        p = lambda a: C.c_void_p(a.ctypes.data)  # noqa: E731
        with torch.cuda.device(self.device):
            self._ck(self.lib.rcv_horn_batch_host(self.h, p(model), 3 * n if model.ndim == 3 else 0, p(est), n, B, p(RT), _stream()))
        return RT


_default = {}


def default_context(device=0, **kw):
    """Process-wide context for the drop-in shim (created on first use)."""
    key = (device, tuple(sorted(kw.items())))
    if key not in _default:
        _default[key] = VoteContext(device, **kw)
    return _default[key]
