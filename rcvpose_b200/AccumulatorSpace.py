"""Drop-in for the voting functions of the reference's AccumulatorSpace.py -- same names, positional
signatures, dtypes, units and return shapes -- executed on the B200 through librcvvote.so.

    rgbd_to_point_cloud(K, depth)     reference AccumulatorSpace.py:77-85
    Accumulator_3D(xyz, radial_list)  reference AccumulatorSpace.py:373-419
    linemod_K                         reference AccumulatorSpace.py:59-61
    read_depth(path)                  reference AccumulatorSpace.py:482-490
    estimate_6d_pose_lm(opts)         reference AccumulatorSpace.py:495-744 (batched: rcvpose_b200/evaluate.py)
    estimate_6d_pose_lmo(opts)        reference AccumulatorSpace.py:741-983 (batched: rcvpose_b200/evaluate.py)
    estimate_6d_pose_ycb(opts)        reference AccumulatorSpace.py:976-1197, repaired (batched: rcvpose_b200/evaluate.py)

Putting this package's directory ahead of the reference on sys.path makes
`estimate_6d_pose_*` (reference :495-1197) call these instead.  Inputs and outputs are NumPy
arrays on the host (the functions synchronise); the batched, device-resident API is
rcvpose_b200.api.VoteContext / rcvpose_b200.pipeline.  No CPU implementation exists here: without
the CUDA library or a GPU the functions raise.
"""
import numpy as np
import torch

try:
    from . import api
except ImportError:
    # Zero-change drop-in: this directory was put on sys.path, so the module was imported
    # top-level as `AccumulatorSpace` (reference train.py:12, AccumulatorSpace.py:2) and has no parent package.
    import os as _os
    import sys as _sys
    _root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    if _root not in _sys.path:
        _sys.path.append(_root)
    from rcvpose_b200 import api

linemod_K = np.array([[572.4114, 0., 325.2611],
                      [0., 573.57043, 242.04899],
                      [0., 0., 1.]])

_CTX_KW = dict(max_items=8, max_points_total=640 * 480 * 2, max_grid=640)


def _ctx():
    return api.default_context(torch.cuda.current_device() if torch.cuda.is_available() else 0, **_CTX_KW)


def rgbd_to_point_cloud(K, depth):
    """(N,3) float64 [x,y,z] of the non-zero depth pixels in row-major order, in depth units."""
    ctx = _ctx()
    depth = np.ascontiguousarray(depth)
    if depth.dtype not in (np.uint16, np.float32, np.float64):
        depth = depth.astype(np.float64)
    d = torch.from_numpy(depth.view(np.int16) if depth.dtype == np.uint16 else depth).to(ctx.device)
    Kd = torch.from_numpy(np.ascontiguousarray(K, dtype=np.float64)).to(ctx.device)
    return ctx.backproject(Kd, d).cpu().numpy()


def Accumulator_3D(xyz, radial_list, acc_unit=5, radius_scale=100, policy=api.RCV_POLICY_LM, return_info=False):
    """xyz (N,3) metres float64, radial_list (N,) decimetres float32/float64 -> (1,3) float64 mm.
    Row 0 of the reference's (M,3) result (callers only use row 0, AccumulatorSpace.py:637).
    Raises ValueError for an empty cloud, as the reference does (xyz_mm.min() on an empty array)."""
    ctx = _ctx()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    r = np.asarray(radial_list)
    r = np.ascontiguousarray(r if r.dtype == np.float32 else r.astype(np.float64))
    if xyz.shape[0] == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    if xyz.shape[0] != r.shape[0]:
        raise ValueError("xyz and radial_list lengths differ")
    out = ctx.vote_points(torch.from_numpy(xyz).to(ctx.device), torch.from_numpy(r).to(ctx.device), acc_unit=acc_unit,
                          radius_scale=radius_scale, policy=policy)
    st = int(out["status"].item())
    if st & api.RCV_ST_EMPTY_MASK:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    if st & api.RCV_ST_BAD_GRID:
        raise ValueError("negative dimensions are not allowed")
    if st != 0:
        raise api.RcvError("Accumulator_3D: status %d (grid %d exceeds capacity?)" % (st, int(out["grid"].item())))
    centre = out["centre_mm"].cpu().numpy().reshape(1, 3)
    if return_info:
        return centre, dict(D=int(out["grid"].item()), zb=int(out["zero_boundary"].item()), peak=int(out["peak"].item()),
                            votes=int(out["votes"].item()))
    return centre


def vote_volume(xyz, radial_list, acc_unit=5, radius_scale=100, policy=api.RCV_POLICY_LM):
    """The reference's VoteMap_3D (AccumulatorSpace.py:399-403) as int32 (D,D,D) -- parity/debug only."""
    ctx = _ctx()
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    r = np.asarray(radial_list)
    r = np.ascontiguousarray(r if r.dtype == np.float32 else r.astype(np.float64))
    x, rr = torch.from_numpy(xyz).to(ctx.device), torch.from_numpy(r).to(ctx.device)
    out = ctx.vote_points(x, rr, acc_unit=acc_unit, radius_scale=radius_scale, policy=policy)
    if int(out["status"].item()) != 0:
        raise api.RcvError("vote_volume: status %d" % int(out["status"].item()))
    D = int(out["grid"].item())
    out = ctx.vote_points(x, rr, acc_unit=acc_unit, radius_scale=radius_scale, policy=policy, want_volume=True, volume_capacity=D ** 3)
    return out["volume_flat"][: D ** 3].reshape(D, D, D).cpu().numpy(), out


def read_depth(path):
    from rcvpose_b200 import formats
    return formats.read_depth(path)


def estimate_6d_pose_lm(opts):
    from rcvpose_b200 import evaluate
    return evaluate.estimate_6d_pose_lm(opts)


def estimate_6d_pose_lmo(opts):
    from rcvpose_b200 import evaluate
    return evaluate.estimate_6d_pose_lmo(opts)


def estimate_6d_pose_ycb(opts):
    from rcvpose_b200 import evaluate
    return evaluate.estimate_6d_pose_ycb(opts)
