"""ctypes front-end for the CPU oracle (oracle/rcv_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of rcv_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module; the product package
``rcvpose_b200`` never does.

The functions mirror the reference's call surface for the voting path:
  rgbd_to_point_cloud   AccumulatorSpace.py:77-85
  Accumulator_3D        AccumulatorSpace.py:373-419   (prelude :373-401, fast_for :325-341, peak :406-419)
  lmshorn               util/horn.py:75-181
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

POLICY_LM = 0      # AccumulatorSpace.py:373-419
POLICY_YCBGEN = 1  # 3DRadius_ycb.py:113-161


def build(force=False):
    so = os.path.join(_HERE, "librcv_oracle.so")
    src = os.path.join(_HERE, "rcv_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        import fcntl
        with open(so + ".lock", "w") as lock:      # ranks / xdist workers may all get here at once
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
                    tmp = "librcv_oracle.tmp.%d.so" % os.getpid()
                    subprocess.run(["make", "-C", _HERE, "-B", tmp], check=True, capture_output=True)
                    os.replace(os.path.join(_HERE, tmp), so)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.orc_pairwise_sum.restype = C.c_double
        L.orc_pairwise_sum.argtypes = [vp, C.c_long, C.c_long]
        L.orc_backproject.restype = C.c_long
        L.orc_backproject.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, vp]
        L.orc_prelude.restype = C.c_int
        L.orc_prelude.argtypes = [vp, C.c_long, vp, C.c_int, C.c_double, C.c_double, C.c_int, vp, vp, vp, ip, ip, dp]
        L.orc_fast_for.restype = None
        L.orc_fast_for.argtypes = [vp, vp, C.c_long, C.c_int, vp, C.c_int]
        L.orc_fast_for_full.restype = None
        L.orc_fast_for_full.argtypes = [vp, vp, C.c_long, C.c_int, vp, C.c_int]
        L.orc_fast_for_literal.restype = None
        L.orc_fast_for_literal.argtypes = [vp, vp, C.c_long, C.c_int, vp]
        L.orc_scatter.restype = None
        L.orc_scatter.argtypes = [vp, vp, C.c_long, C.c_int, vp]
        L.orc_peak.restype = C.c_long
        L.orc_peak.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_int32)]
        L.orc_center_mm.restype = None
        L.orc_center_mm.argtypes = [vp, C.c_int, vp, C.c_double, C.c_int, vp]
        L.orc_lmshorn.restype = None
        L.orc_lmshorn.argtypes = [vp, vp, C.c_int, vp]
        L.orc_accumulator_3d.restype = C.c_int
        L.orc_accumulator_3d.argtypes = [vp, C.c_long, vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                         vp, ip, ip, C.POINTER(C.c_int32), C.POINTER(C.c_longlong), vp]
        L.orc_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().orc_num_threads()


def pairwise_sum(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return lib().orc_pairwise_sum(_p(a), a.size, 1)


def rgbd_to_point_cloud(K, depth, swap_xy=False):
    """AccumulatorSpace.py:77-85 -> (N,3) float64 in depth units."""
    K = np.ascontiguousarray(K, dtype=np.float64)
    d = np.ascontiguousarray(depth, dtype=np.float64)
    H, W = d.shape
    n = lib().orc_backproject(_p(K), _p(d), H, W, int(swap_xy), None)
    out = np.empty((n, 3), dtype=np.float64)
    lib().orc_backproject(_p(K), _p(d), H, W, int(swap_xy), _p(out))
    return out


def _radii(radial_list):
    r = np.asarray(radial_list)
    if r.dtype == np.float32:
        return np.ascontiguousarray(r), 1
    return np.ascontiguousarray(r, dtype=np.float64), 0


def prelude(xyz, radial_list, acc_unit=5.0, radius_scale=100.0, policy=POLICY_LM):
    """AccumulatorSpace.py:373-401 -> dict(p (n,3) f64, R (n,) i32, mean (3,), zb, D, rmax)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    n = xyz.shape[0]
    if n == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    r, is32 = _radii(radial_list)
    p = np.empty((n, 3), dtype=np.float64)
    R = np.empty(n, dtype=np.int32)
    mean = np.empty(3, dtype=np.float64)
    zb, D, rmax = C.c_int(), C.c_int(), C.c_double()
    lib().orc_prelude(_p(xyz), n, _p(r), is32, acc_unit, radius_scale, policy, _p(p), _p(R), _p(mean),
                      C.byref(zb), C.byref(D), C.byref(rmax))
    return dict(p=p, R=R, mean=mean, zb=zb.value, D=D.value, rmax=rmax.value)


def fast_for(p, R, D, threads=0, method="brute"):
    """AccumulatorSpace.py:325-341 on voxel-unit points with integer radii -> int32 (D,D,D)."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.int32)
    vol = np.zeros((D, D, D), dtype=np.int32)
    if method == "brute":
        lib().orc_fast_for(_p(p), _p(R), p.shape[0], D, _p(vol), threads)
    elif method == "full":
        lib().orc_fast_for_full(_p(p), _p(R), p.shape[0], D, _p(vol), threads)
    elif method == "literal":
        lib().orc_fast_for_literal(_p(p), _p(R), p.shape[0], D, _p(vol))
    elif method == "scatter":
        lib().orc_scatter(_p(p), _p(R), p.shape[0], D, _p(vol))
    else:
        raise ValueError(method)
    return vol


def peak(vol):
    """AccumulatorSpace.py:406 -> (first argmax (i,j,k) in C order, max count, number of ties)."""
    vol = np.ascontiguousarray(vol, dtype=np.int32)
    idx = np.empty(3, dtype=np.int32)
    mx = C.c_int32()
    ties = lib().orc_peak(_p(vol), vol.shape[0], _p(idx), C.byref(mx))
    return idx, mx.value, ties


def center_mm(idx, zb, mean, acc_unit=5.0, policy=POLICY_LM):
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    out = np.empty(3, dtype=np.float64)
    lib().orc_center_mm(_p(idx), zb, _p(mean), acc_unit, policy, _p(out))
    return out


def Accumulator_3D(xyz, radial_list, acc_unit=5.0, radius_scale=100.0, policy=POLICY_LM, method="brute", threads=0,
                   return_info=False, return_volume=False):
    """AccumulatorSpace.py:373-419.  Returns (1,3) float64 mm (row 0 of the reference's result)."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    n = xyz.shape[0]
    if n == 0:
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    r, is32 = _radii(radial_list)
    c = np.empty(3, dtype=np.float64)
    D, zb, pk, votes = C.c_int(), C.c_int(), C.c_int32(), C.c_longlong()
    vol = None
    if return_volume:
        pre = prelude(xyz, r, acc_unit, radius_scale, policy)
        vol = np.zeros((pre["D"],) * 3, dtype=np.int32)
    rc = lib().orc_accumulator_3d(_p(xyz), n, _p(r), is32, acc_unit, radius_scale, policy, {"brute": 1, "full": 2}.get(method, 0), threads,
                                  _p(c), C.byref(D), C.byref(zb), C.byref(pk), C.byref(votes),
                                  _p(vol) if vol is not None else None)
    if rc == 2:
        raise ValueError("negative dimensions are not allowed")
    out = c.reshape(1, 3)
    if return_info or return_volume:
        info = dict(D=D.value, zb=zb.value, peak=pk.value, votes=votes.value)
        if return_volume:
            info["volume"] = vol
        return out, info
    return out


def lmshorn(P1, P2, n, A):
    """util/horn.py:75-181: fills the 4x4 A in place; P1/P2 untouched."""
    P1 = np.ascontiguousarray(P1, dtype=np.float64)
    P2 = np.ascontiguousarray(P2, dtype=np.float64)
    out = np.empty((4, 4), dtype=np.float64)
    lib().orc_lmshorn(_p(P1), _p(P2), int(n), _p(out))
    A[...] = out


def add_metric(model_mm, RT_est, RT_gt):
    """ADD(-S) distance before ICP as the reference computes it (AccumulatorSpace.py:664-702): the CAD points under the
    estimated and the ground-truth pose (project(), :64-75), then for every ground-truth point the distance to the nearest
    estimated point (open3d compute_point_cloud_distance = exact nearest neighbour; a k-d tree here).  Returns (mean, min)."""
    from scipy.spatial import cKDTree
    est = np.dot(model_mm, RT_est[:3, :3].T) + RT_est[:3, 3:].T
    gt = np.dot(model_mm, RT_gt[:3, :3].T) + RT_gt[:3, 3:].T
    d, _ = cKDTree(est).query(gt, k=1)
    return float(d.mean()), float(d.min())


def scene_union(clouds):
    """xyz_mm_icp (AccumulatorSpace.py:620-625, :863-868, :1070-1075): the first keypoint's cloud, then every point of the later
    clouds that is not in the list yet, in order of first appearance (the reference's O(N^2) loop, restated with a set of rows)."""
    out = [np.asarray(c, dtype=np.float64) for c in clouds[:1]]
    seen = set(map(bytes, out[0])) if out else set()
    for c in clouds[1:]:
        keep = []
        for row in np.asarray(c, dtype=np.float64):
            b = bytes(row)
            if b not in seen:
                seen.add(b)
                keep.append(row)
        if keep:
            out.append(np.array(keep))
    return np.concatenate(out, axis=0) if out else np.zeros((0, 3))


def umeyama_rigid(src, dst):
    """Eigen::umeyama(src, dst, with_scaling=false) on (n,3) arrays -> 4x4: the rigid transform of
    TransformationEstimationPointToPoint::ComputeTransformation (open3d 0.14.1).  Restated from Eigen/src/Geometry/Umeyama.h:
    sigma = dst_demean * src_demean^T / n, SVD, S = diag(1, 1, sign(det U det V)), R = U S V^T, t = mean_dst - R mean_src."""
    n = src.shape[0]
    ms, md = src.mean(axis=0), dst.mean(axis=0)
    sigma = (dst - md).T @ (src - ms) / n
    U, _, Vt = np.linalg.svd(sigma)
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1.0
    R = U @ np.diag(S) @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = md - R @ ms
    return T


def registration_icp(source, target, max_correspondence_distance, init, max_iteration=30, relative_fitness=1e-6, relative_rmse=1e-6):
    """open3d.pipelines.registration.registration_icp(source, target, max_correspondence_distance, init,
    TransformationEstimationPointToPoint(), ICPConvergenceCriteria(...)) as the reference calls it
    (AccumulatorSpace.py:704-718, :929-950, :1152-1180).  open3d 0.14.1 (rcvpose.yml:176) is a third-party dependency that is
    not in the reference tree nor in this image, so this is a restatement of its published algorithm
    (cpp/open3d/pipelines/registration/Registration.cpp: RegistrationICP + GetRegistrationResultAndCorrespondences) and
    parity for this row is UNPINNED (no open3d output to check against).  Returns dict(transformation, fitness, inlier_rmse,
    iterations)."""
    from scipy.spatial import cKDTree
    source = np.asarray(source, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    tree = cKDTree(target) if len(target) else None

    def evaluate(pcd):
        if tree is None:
            return 0.0, 0.0, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        d, j = tree.query(pcd, k=1)
        keep = d * d < max_correspondence_distance * max_correspondence_distance    # SearchHybrid: strict
        i = np.nonzero(keep)[0]
        if len(i) == 0:
            return 0.0, 0.0, i, j[i]
        return len(i) / len(pcd), float(np.sqrt(np.sum(d[i] ** 2) / len(i))), i, j[i]

    T = np.array(init, dtype=np.float64)
    pcd = source @ T[:3, :3].T + T[:3, 3]
    fit, rmse, ci, cj = evaluate(pcd)
    it = 0
    for it in range(1, max_iteration + 1):
        upd = umeyama_rigid(pcd[ci], target[cj]) if len(ci) else np.eye(4)
        T = upd @ T
        pcd = pcd @ upd[:3, :3].T + upd[:3, 3]
        bfit, brmse = fit, rmse
        fit, rmse, ci, cj = evaluate(pcd)
        if abs(bfit - fit) < relative_fitness and abs(brmse - rmse) < relative_rmse:
            break
    return dict(transformation=T, fitness=fit, inlier_rmse=rmse, iterations=it)
