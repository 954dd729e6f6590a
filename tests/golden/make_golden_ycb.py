"""Golden fixture for the YCB-Video evaluator: runs the reference's own `estimate_6d_pose_ycb`
(AccumulatorSpace.py:976-1197, imported from /root/reference) on a synthetic class laid out by
rcvpose_b200.synth.write_ycb_dataset and stores what it computed.

The function cannot run as written.  It is REPAIRED IN MEMORY, by textual substitution on
`inspect.getsource` at generation time -- nothing of the reference is copied into this repository:
    :1003  model_path built from `keypoint_count` before assignment      -> from the loop variable i
    :1016  RTGT = poses[:, :, np.where(cls_indexes == class_id)]          -> the pose of the object whose cls_index is class_id
    :1044  model_list[keypoint_count]   (1..3 into a 3-element list)      -> model_list[keypoint_count - 1]
    :1094  estimated_kpts[keypoint_count]  (1..3 into a (3,3) array)      -> estimated_kpts[keypoint_count - 1]
    :1115  kpts = keypoints[0:3]  (estimates are of keypoints 1..3)       -> keypoints[1:4]
Stand-ins, as in make_golden_evaluator.py: open3d (points.xyz reader, nearest-neighbour distances, the oracle's ICP, an
oriented bounding box from the PCA of the points), h5py (a file object whose 'JPEGImages/' keys are the frames on disk), the
three networks (map files written by synth.write_ycb_ckpt_maps).  Everything else -- the frame loop, loadmat, the mask rule,
depth / factor_depth, rgbd_to_point_cloud with the frame's intrinsics, Accumulator_3D, lmshorn, project, thresholds, AUC
counters and the print-out -- is the reference's code.

Usage (build container only):  NUMBA_NUM_THREADS=1 python tests/golden/make_golden_ycb.py
"""
import inspect
import os
import re
import sys
import tempfile
import types

os.environ.setdefault("NUMBA_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402

import make_golden_evaluator as G  # noqa: E402  (imports the reference with the open3d / h5py / matplotlib stand-ins)
from rcvpose_b200 import evaluate, formats, synth  # noqa: E402

A = G.A
CLASSES = [(5, 3, 21), (13, 2, 22)]        # (class id, frames, seed): 006_mustard_bottle, 024_bowl (symmetric: ADD-S)

REPAIRS = [
    ('model_path = opts.model_dir + class_name+"_pt"+str(keypoint_count)+".pth.tar"', 'model_path = opts.model_dir + class_name+"_pt"+str(i)+".pth.tar"'),
    ("RTGT = sceneInfo['poses'][:,:,np.where(sceneInfo['cls_indexes']==class_id)]", "RTGT = sceneInfo['poses'][:,:,int(np.where(np.ravel(sceneInfo['cls_indexes'])==class_id)[0][0])]"),
    ("FCResBackbone(model_list[keypoint_count], input_path, normalized_depth)", "FCResBackbone(model_list[keypoint_count-1], input_path, normalized_depth)"),
    ("estimated_kpts[keypoint_count] = estimated_center_mm", "estimated_kpts[keypoint_count-1] = estimated_center_mm"),
    ("kpts = keypoints[0:3,:]*1000", "kpts = keypoints[1:4,:]*1000"),
]


def repaired_ycb():
    src = inspect.getsource(A.estimate_6d_pose_ycb)
    for old, new in REPAIRS:
        assert src.count(old) == 1, old
        src = src.replace(old, new)
    ns = {}
    exec(compile(src, "<estimate_6d_pose_ycb, repaired in memory>", "exec"), A.__dict__, ns)
    return ns["estimate_6d_pose_ycb"]


class _Cloud(G._Cloud):
    def get_oriented_bounding_box(self):
        p = np.asarray(self.points, dtype=np.float64)
        c = p - p.mean(0)
        _, vec = np.linalg.eigh(c.T @ c / len(p))
        q = c @ vec
        return types.SimpleNamespace(extent=q.max(0) - q.min(0))


class _H5File(dict):
    def __init__(self, path, *a):
        root, cls = os.path.dirname(path.rstrip("/")) + "/", os.path.splitext(os.path.basename(path))[0]
        names = {}
        for cycle in sorted(os.listdir(G.RECORD["root"] + "data/")):
            for f in sorted(os.listdir(G.RECORD["root"] + "data/" + cycle)):
                if f.endswith(".mat"):
                    names[cycle + "_" + os.path.splitext(f)[0]] = None
        super().__init__({"JPEGImages/": names})


def _fake_backbone(model, input_img_path, normalized_depth):
    cycle = os.path.basename(os.path.dirname(input_img_path))
    idx = os.path.basename(input_img_path).split("-")[0]
    G.RECORD["frames"].append(cycle + "_" + idx)
    return synth.load_ycb_ckpt_maps(G.RECORD["root"], G.RECORD["cls"], model.k, cycle + "_" + idx)


def main():
    import torch
    o3d = sys.modules["open3d"]
    o3d.io = types.SimpleNamespace(read_point_cloud=lambda p: _Cloud(formats.read_xyz_points(p) if p.endswith(".xyz") else formats.read_ply_points(p)))
    o3d.geometry = types.SimpleNamespace(PointCloud=_Cloud)
    sys.modules["h5py"].File = _H5File
    A.h5py = sys.modules["h5py"]
    A.DenseFCNResNet152 = G._FakeNet
    A.utils.load_checkpoint = lambda model, optim, path: (model, optim, 0, 0)
    A.FCResBackbone = _fake_backbone
    A.torch.nn.DataParallel = lambda m: m
    fn = repaired_ycb()
    out = {}
    for class_id, n_frames, seed in CLASSES:
        cls = evaluate.ycb_cls_names[class_id]
        root = tempfile.mkdtemp() + "/"
        names = synth.write_ycb_dataset(root, class_id, cls, n_frames, seed=seed)
        synth.write_ycb_ckpt_maps(root, class_id, cls, names, seed=seed)
        os.makedirs(root + "ckpts")
        for k in (1, 2, 3):
            open(root + "ckpts/" + cls + "_pt" + str(k) + ".pth.tar", "wb").close()
        A.ycb_cls_names = {class_id: cls}
        G._FakeNet.count = 0
        G.RECORD["root"], G.RECORD["cls"], G.RECORD["frames"] = root, cls, []
        text = G.run(fn, types.SimpleNamespace(root_dataset=root, model_dir=root + "ckpts/", demo_mode=False))
        frames = G.RECORD["frames"][::3]
        assert frames == sorted(names), (frames, names)
        tag = "ycb_%d" % class_id
        sym = cls in A.ycb_syms
        d = np.array(G.RECORD["distances"]).reshape(n_frames, -1, 2)       # per frame: 2 calls before ICP, 2 after
        out[tag + "_frames"] = np.array(frames)
        out[tag + "_centres"] = np.array(G.RECORD["centres"]).reshape(n_frames, 3, 3)
        out[tag + "_n_points"] = np.array(G.RECORD["n_points"]).reshape(n_frames, 3)
        out[tag + "_RT"] = np.array(G.RECORD["RT"])
        out[tag + "_dist_before"] = d[:, 0, 1 if sym else 0]
        out[tag + "_mean_before"] = d[:, 0, 0]
        out[tag + "_dist_after"] = d[:, -1, 1 if sym else 0]
        out[tag + "_icp_RT"] = np.array([r[0] for r in G.RECORD["icp"]])
        out[tag + "_icp_iters"] = np.array([r[1] for r in G.RECORD["icp"]])
        out[tag + "_icp_threshold"] = np.array([r[3] for r in G.RECORD["icp"]])
        out[tag + "_scene_points"] = np.array([r[4] for r in G.RECORD["icp"]])
        nums = {}
        for key, pat in (("auc_before", r"AUC of " + cls + r" before ICP:\s+(\S+)"), ("auc_after", r"AUC of " + cls + r" after ICP:\s+(\S+)"),
                         ("add_before", r"^ADD\\\(s\\\) of " + cls + r" before ICP:\s+(\S+)"), ("add_after", r"^ADD\\\(s\\\) of " + cls + r" after ICP:\s+(\S+)")):
            nums[key] = float(re.findall(pat, text, flags=re.M)[-1])
        out[tag + "_summary"] = np.array([nums["auc_before"], nums["auc_after"], nums["add_before"], nums["add_after"]])
        out[tag + "_seed"] = np.array([n_frames, seed])
        print(tag, frames, out[tag + "_summary"], out[tag + "_n_points"].tolist(), out[tag + "_dist_before"], out[tag + "_dist_after"], out[tag + "_icp_iters"])
    np.savez_compressed(os.path.join(HERE, "ycb_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "ycb_golden.npz"))


if __name__ == "__main__":
    main()
