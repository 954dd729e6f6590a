"""Golden fixture for the evaluator rows: runs the REAL reference's `estimate_6d_pose_lm` and
`estimate_6d_pose_lmo` (aaronWool/rcvpose, unmodified, imported from /root/reference) in this container on small synthetic
datasets laid out by rcvpose_b200.synth.write_lm_dataset / write_lmo_dataset, and stores what it computed.

Usage (build container only -- /root/reference does not exist on the GPU box):
    NUMBA_NUM_THREADS=1 python tests/golden/make_golden_evaluator.py

The reference's evaluator needs open3d 0.14.1 (rcvpose.yml:176), which is neither in the reference tree nor in this image.  A
stand-in module provides exactly the calls the evaluator makes: read_point_cloud (our PLY reader), PointCloud with
compute_point_cloud_distance (exact nearest neighbour, k-d tree) and transform, Vector3dVector, and registration_icp (the
oracle's restatement of open3d's published algorithm).  Everything else is the reference's own code running on its own
control flow: directory walking, `read_depth`, max_radii, the mask rule, `rgbd_to_point_cloud`, `Accumulator_3D`, the
append-if-new union, `lmshorn`, `project`, the ADD(-S) thresholds, counters and print-out.  So the fixture PINS the evaluator
up to and including the ADD(-S) distance before ICP and the counters; the refined pose and the distance after ICP depend on the
stand-in's ICP and stay "unpinned" (DESIGN.md section 2).  Only outputs are stored; no reference source is copied.
"""
import contextlib
import io
import os
import re
import sys
import tempfile
import types

os.environ.setdefault("NUMBA_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from scipy.spatial import cKDTree  # noqa: E402

from oracle import oracle  # noqa: E402
from rcvpose_b200 import formats, synth  # noqa: E402

LM_CLASSES = [("ape", 3, 11, np.float32), ("eggbox", 2, 12, np.float32)]      # (class, frames, seed, radius map dtype)
LMO_CLASSES = [("can", 3, 6, np.float32)]
LM_CKPT_CLASSES = [("cat", 3, 13)]                                              # checkpoint branch: (class, frames, seed)


class _Cloud:
    def __init__(self, points=None):
        self.points = np.zeros((0, 3)) if points is None else np.asarray(points, dtype=np.float64)

    def paint_uniform_color(self, c):
        return self

    def compute_point_cloud_distance(self, other):
        a, b = np.asarray(self.points, dtype=np.float64), np.asarray(other.points, dtype=np.float64)
        d = cKDTree(b).query(a, k=1)[0] if len(a) and len(b) else np.zeros(0)
        RECORD["distances"].append((float(d.mean()) if d.size else np.nan, float(d.min()) if d.size else np.nan))
        return d

    def transform(self, T):
        T = np.asarray(T, dtype=np.float64)
        self.points = np.asarray(self.points, dtype=np.float64) @ T[:3, :3].T + T[:3, 3]
        return self


class _Criteria:
    def __init__(self, relative_fitness=1e-6, relative_rmse=1e-6, max_iteration=30):
        self.relative_fitness, self.relative_rmse, self.max_iteration = relative_fitness, relative_rmse, max_iteration


def _registration_icp(source, target, threshold, init, estimation=None, criteria=None):
    criteria = criteria or _Criteria()
    reg = oracle.registration_icp(source.points, target.points, threshold, init, max_iteration=criteria.max_iteration,
                                  relative_fitness=criteria.relative_fitness, relative_rmse=criteria.relative_rmse)
    RECORD["icp"].append((np.array(reg["transformation"]), reg["iterations"], reg["fitness"], float(threshold), len(np.asarray(target.points))))
    return types.SimpleNamespace(transformation=reg["transformation"], fitness=reg["fitness"], inlier_rmse=reg["inlier_rmse"])


def _fake_open3d():
    o3d = types.ModuleType("open3d")
    o3d.io = types.SimpleNamespace(read_point_cloud=lambda p: _Cloud(formats.read_ply_points(p)))
    o3d.geometry = types.SimpleNamespace(PointCloud=_Cloud)
    o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: np.asarray(a, dtype=np.float64))
    o3d.visualization = types.SimpleNamespace(draw_geometries=lambda *a, **k: None)
    o3d.pipelines = types.SimpleNamespace(registration=types.SimpleNamespace(
        ICPConvergenceCriteria=_Criteria, registration_icp=_registration_icp, TransformationEstimationPointToPoint=lambda: None))
    return o3d


RECORD = {}
sys.modules["open3d"] = _fake_open3d()
for name in ("h5py", "matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")
import numba  # noqa: E402
import AccumulatorSpace as A  # noqa: E402

assert numba.get_num_threads() == 1, "run with NUMBA_NUM_THREADS=1"

_acc3d, _read_depth, _Horn = A.Accumulator_3D, A.read_depth, A.HornPoseFitting


def _rec_acc3d(xyz, radial_list):
    out = _acc3d(xyz, radial_list)
    RECORD["centres"].append(np.array(out[0], dtype=np.float64))
    RECORD["n_points"].append(len(radial_list))
    return out


def _rec_read_depth(path):
    RECORD["depth_paths"].append(path)
    return _read_depth(path)


class _RecHorn(_Horn):
    def lmshorn(self, P1, P2, n, RT):
        RECORD["est_kpts"].append(np.array(P2, dtype=np.float64))     # before the call: lmshorn centres and restores P2 in place (1 ulp)
        super().lmshorn(P1, P2, n, RT)
        RECORD["RT"].append(np.array(RT, dtype=np.float64))


A.Accumulator_3D, A.read_depth, A.HornPoseFitting = _rec_acc3d, _rec_read_depth, _RecHorn


# ---- checkpoint branch: the networks are replaced by map files (synth.write_lm_ckpt_maps); the reference's own code decides what
# to do with the maps (AccumulatorSpace.py:594-610) ----
class _FakeNet:
    count = 0

    def __init__(self, *a):
        _FakeNet.count += 1
        self.k = (_FakeNet.count - 1) % 3 + 1          # the reference builds the three keypoint networks in order (:518-527)

    def parameters(self):
        import torch
        return [torch.nn.Parameter(torch.zeros(1))]

    def eval(self):
        return self


def _fake_backbone(model, input_img_path, normalized_depth):
    stem = os.path.splitext(os.path.basename(input_img_path))[0]
    return synth.load_lm_ckpt_maps(RECORD["root"], RECORD["cls"], model.k, stem)


def run(fn, opts):
    for k in ("centres", "n_points", "depth_paths", "RT", "est_kpts", "distances", "icp"):
        RECORD[k] = []
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(opts)
    return buf.getvalue()


def final_ratios(text, cls):
    b = re.findall(r"^ADDs? of " + cls + r" before ICP:\s+(\S+)", text, flags=re.M)
    a = re.findall(r"^ADDs? of " + cls + r" after ICP:\s+(\S+)", text, flags=re.M)
    return float(b[-1]), float(a[-1])


def main():
    out = {}
    # ---- LINEMOD ----
    for cls, n_frames, seed, dt in LM_CLASSES:
        root = tempfile.mkdtemp() + "/"
        synth.write_lm_dataset(root, cls, n_frames, seed=seed, radius_dtype=dt)
        A.lm_cls_names = [cls]
        text = run(A.estimate_6d_pose_lm, types.SimpleNamespace(root_dataset=root, model_dir="", using_ckpts=False, demo_mode=False))
        stems = ["%06d" % int(re.search(r"depth(\d+)\.dpt$", p).group(1)) for p in RECORD["depth_paths"][::3]]
        order = np.argsort(stems)
        tag = "lm_" + cls
        out[tag + "_stems"] = np.array(stems)[order]
        out[tag + "_centres"] = np.array(RECORD["centres"]).reshape(n_frames, 3, 3)[order]
        out[tag + "_n_points"] = np.array(RECORD["n_points"]).reshape(n_frames, 3)[order]
        out[tag + "_RT"] = np.array(RECORD["RT"])[order]
        d = np.array(RECORD["distances"])            # two calls before ICP (threshold test + value), two after, per frame
        per = len(d) // n_frames
        d = d.reshape(n_frames, per, 2)[order]
        sym = cls in A.lm_syms
        out[tag + "_dist_before"] = d[:, 0, 1 if sym else 0]
        out[tag + "_dist_after"] = d[:, -1, 1 if sym else 0]
        out[tag + "_icp_RT"] = np.array([r[0] for r in RECORD["icp"]])[order]
        out[tag + "_icp_iters"] = np.array([r[1] for r in RECORD["icp"]])[order]
        out[tag + "_icp_threshold"] = np.array([r[3] for r in RECORD["icp"]])[order]
        out[tag + "_scene_points"] = np.array([r[4] for r in RECORD["icp"]])[order]
        out[tag + "_ratios"] = np.array(final_ratios(text, cls))
        out[tag + "_seed"] = np.array([n_frames, seed])
        print(tag, out[tag + "_stems"], out[tag + "_ratios"], out[tag + "_dist_before"], out[tag + "_dist_after"], out[tag + "_icp_iters"])
    # ---- Occlusion LINEMOD ----
    for cls, n_frames, seed, dt in LMO_CLASSES:
        root = tempfile.mkdtemp() + "/"
        synth.write_lmo_dataset(root, cls, n_frames, seed=seed, radius_dtype=dt)
        A.lmo_cls_names = [cls]
        text = run(A.estimate_6d_pose_lmo, types.SimpleNamespace(root_dataset=root, model_dir="", using_ckpts=False, demo_mode=False))
        tag = "lmo_" + cls
        out[tag + "_RT"] = np.array(RECORD["RT"])                     # in os.listdir order of the evaluated frames
        out[tag + "_est_kpts"] = np.array(RECORD["est_kpts"])
        out[tag + "_icp_RT"] = np.array([r[0] for r in RECORD["icp"]])
        out[tag + "_icp_iters"] = np.array([r[1] for r in RECORD["icp"]])
        out[tag + "_scene_points"] = np.array([r[4] for r in RECORD["icp"]])
        out[tag + "_ratios"] = np.array(final_ratios(text, cls))
        out[tag + "_seed"] = np.array([n_frames, seed])
        print(tag, out[tag + "_ratios"], out[tag + "_icp_iters"], out[tag + "_est_kpts"][:, :, 0])
    # ---- LINEMOD, checkpoint branch ----
    A.DenseFCNResNet152 = _FakeNet
    A.utils.load_checkpoint = lambda model, optim, path: (model, optim, 0, 0)
    A.FCResBackbone = _fake_backbone
    for cls, n_frames, seed in LM_CKPT_CLASSES:
        root = tempfile.mkdtemp() + "/"
        stems = synth.write_lm_dataset(root, cls, n_frames, seed=seed)
        synth.write_lm_ckpt_maps(root, cls, stems, seed=seed)
        os.makedirs(root + "ckpts")
        for k in (1, 2, 3):
            open(root + "ckpts/" + cls + "_pt" + str(k) + ".pth.tar", "wb").close()
        A.lm_cls_names = [cls]
        _FakeNet.count = 0
        RECORD["root"], RECORD["cls"] = root, cls
        text = run(A.estimate_6d_pose_lm, types.SimpleNamespace(root_dataset=root, model_dir=root + "ckpts/", using_ckpts=True, demo_mode=False))
        stems_run = ["%06d" % int(re.search(r"depth(\d+)\.dpt$", p).group(1)) for p in RECORD["depth_paths"][::3]]
        order = np.argsort(stems_run)
        tag = "lmckpt_" + cls
        out[tag + "_stems"] = np.array(stems_run)[order]
        out[tag + "_centres"] = np.array(RECORD["centres"]).reshape(n_frames, 3, 3)[order]
        out[tag + "_n_points"] = np.array(RECORD["n_points"]).reshape(n_frames, 3)[order]
        out[tag + "_RT"] = np.array(RECORD["RT"])[order]
        d = np.array(RECORD["distances"])
        d = d.reshape(n_frames, len(d) // n_frames, 2)[order]
        out[tag + "_dist_before"] = d[:, 0, 0]
        out[tag + "_scene_points"] = np.array([r[4] for r in RECORD["icp"]])[order]
        out[tag + "_ratios"] = np.array(final_ratios(text, cls))
        out[tag + "_seed"] = np.array([n_frames, seed])
        print(tag, out[tag + "_stems"], out[tag + "_ratios"], out[tag + "_n_points"], out[tag + "_dist_before"])
    np.savez_compressed(os.path.join(HERE, "evaluator_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "evaluator_golden.npz"))


if __name__ == "__main__":
    main()
