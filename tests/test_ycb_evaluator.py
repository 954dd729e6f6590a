"""YCB-Video evaluator (`estimate_6d_pose_ycb`, AccumulatorSpace.py:976-1197).

The fixture tests/golden/ycb_golden.npz holds what the reference's own function computed in the build container after the
in-memory repairs listed in tests/golden/make_golden_ycb.py (the function cannot run as written).  Here:
  * CPU: the per-frame loop restated on the oracle's functions reproduces the fixture (keypoints bit-identical);
  * GPU: rcvpose_b200.evaluate.estimate_6d_pose_ycb reproduces it through librcvvote.so."""
import types

import numpy as np
import pytest
import torch

from oracle import oracle
from rcvpose_b200 import evaluate, formats, synth
from tests.conftest import ROOT  # noqa: F401

CASES = [(5, False), (13, True)]


@pytest.fixture(scope="module")
def ycb_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ycb_golden.npz"))


def _dataset(tmp_path, g, class_id):
    n_frames, seed = (int(v) for v in g["ycb_%d_seed" % class_id])
    root = str(tmp_path) + "/"
    name = evaluate.ycb_cls_names[class_id]
    names = synth.write_ycb_dataset(root, class_id, name, n_frames, seed=seed)
    synth.write_ycb_ckpt_maps(root, class_id, name, names, seed=seed)
    return root, name, n_frames


def _reference_loop_ycb(root, class_id):
    """The repaired per-frame loop (:1012-1191) on the oracle's functions."""
    cls = evaluate.YcbClass(root, class_id)
    sym = cls.name in evaluate.ycb_syms
    thr_mm = evaluate.obb_diagonal(cls.cad_m) * 0.01 * 1000
    rows = []
    for stem in cls.stems:
        meta = cls.meta(stem)
        K, factor = meta["intrinsic_matrix"], meta["factor_depth"]
        depth1 = cls.depth_raw(stem)
        est, clouds, npts = np.zeros((3, 3)), [], []
        for k in (1, 2, 3):
            sem, radial = synth.load_ycb_ckpt_maps(root, cls.name, k, stem)
            m = np.where(sem > 0.8, 1, 0)
            dm = depth1 * m / factor
            xyz = oracle.rgbd_to_point_cloud(K, dm)
            rl = radial[dm.nonzero()]                                 # surviving pixel = mask and depth != 0 for both lists
            est[k - 1] = oracle.Accumulator_3D(xyz, rl)[0]
            clouds.append(xyz * 1000)
            npts.append(len(rl))
        RT = np.zeros((4, 4))
        oracle.lmshorn(cls.keypoints_m[1:4] * 1000, est, 3, RT)
        gt = np.eye(4)
        gt[:3] = formats.pose_of(meta, class_id)
        gt[:3, 3] *= 1000
        mean, mn = oracle.add_metric(cls.cad_m * 1000, RT, gt)
        before = mn if sym else mean
        reg = oracle.registration_icp(cls.cad_m * 1000, oracle.scene_union(clouds), mean, RT, max_iteration=2000000)
        mean2, mn2 = oracle.add_metric(cls.cad_m * 1000, reg["transformation"], gt)
        after = mn2 if sym else mean2
        rows.append(dict(centres=est, n_points=npts, RT=RT, before=before, mean_before=mean, after=after, RT_icp=reg["transformation"],
                         iters=reg["iterations"], pb=before <= thr_mm, pa=after <= thr_mm))
    return cls, rows, thr_mm


def _summary(rows, n):
    thr = np.array(evaluate.ycb_auc_thresholds_m) * 1000
    cb = [sum(r["before"] <= t for r in rows) / n for t in thr]
    ca = [sum(r["after"] <= t for r in rows) / n for t in thr]
    return [evaluate.trapezoid_auc(evaluate.ycb_auc_thresholds_m, cb) / 0.1, evaluate.trapezoid_auc(evaluate.ycb_auc_thresholds_m, ca) / 0.1,
            sum(r["pb"] for r in rows) / n, sum(r["pa"] for r in rows) / n]


def test_ycb_layout_obb_and_auc(tmp_path):
    root = str(tmp_path) + "/"
    names = synth.write_ycb_dataset(root, 5, evaluate.ycb_cls_names[5], 2, seed=1, split_extra=2)
    cls = evaluate.YcbClass(root, 5)
    assert cls.stems == sorted(names) and len(names) == 2            # frames on disk but not in val.txt are not evaluated
    meta = cls.meta(cls.stems[1])
    assert meta["factor_depth"] == 10000.0 and meta["intrinsic_matrix"][0, 2] != synth.ycb_K[0, 2]     # per-frame intrinsics
    assert formats.pose_of(meta, 5) is not None and formats.pose_of(meta, 9) is None
    assert cls.depth_raw(cls.stems[0]).dtype == np.uint16
    # OBB of a box-shaped cloud = the box, whatever its orientation
    rng = np.random.default_rng(0)
    box = rng.uniform(-1, 1, size=(4000, 3)) * np.array([0.3, 0.2, 0.1])
    box = np.concatenate([box, np.array([[0.3, 0.2, 0.1], [-0.3, -0.2, -0.1], [0.3, -0.2, 0.1], [-0.3, 0.2, -0.1]])])
    a = rng.normal(size=3); th = np.linalg.norm(a); k = a / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    want = 2 * np.sqrt(0.3 ** 2 + 0.2 ** 2 + 0.1 ** 2)
    assert abs(evaluate.obb_diagonal(box @ R.T + 5.0) - want) < 0.06 * want
    from sklearn import metrics
    y = [0.0, 0.25, 0.5, 0.5, 1.0, 1.0]
    assert evaluate.trapezoid_auc(evaluate.ycb_auc_thresholds_m, y) == pytest.approx(metrics.auc(evaluate.ycb_auc_thresholds_m, y), abs=1e-15)


@pytest.mark.parametrize("class_id,sym", CASES)
def test_oracle_loop_matches_the_repaired_reference_ycb(tmp_path, ycb_golden, class_id, sym):
    g, tag = ycb_golden, "ycb_%d" % class_id
    root, name, n = _dataset(tmp_path, g, class_id)
    cls, rows, thr = _reference_loop_ycb(root, class_id)
    assert cls.stems == list(g[tag + "_frames"]) and (name in evaluate.ycb_syms) == sym
    for i, r in enumerate(rows):
        assert np.array_equal(r["centres"], g[tag + "_centres"][i])            # Accumulator_3D outputs, bit for bit
        assert r["n_points"] == list(g[tag + "_n_points"][i])
        np.testing.assert_allclose(r["RT"], g[tag + "_RT"][i], rtol=0, atol=1e-9)
        np.testing.assert_allclose(r["before"], g[tag + "_dist_before"][i], rtol=1e-9)
        np.testing.assert_allclose(r["mean_before"], g[tag + "_icp_threshold"][i], rtol=1e-9)   # ICP threshold = MEAN distance (:1151)
        assert r["iters"] == g[tag + "_icp_iters"][i]
        np.testing.assert_allclose(r["RT_icp"], g[tag + "_icp_RT"][i], rtol=0, atol=1e-7)
        np.testing.assert_allclose(r["after"], g[tag + "_dist_after"][i], rtol=1e-6)
    np.testing.assert_allclose(_summary(rows, n), g[tag + "_summary"], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("class_id,sym", CASES)
def test_estimate_6d_pose_ycb_vs_repaired_reference_golden(tmp_path, ycb_golden, class_id, sym, capsys):
    """GPU: the drop-in against what the (repaired) reference computed on the same dataset: keypoints bit-identical, poses,
    ADD(-S) before ICP, the ICP target size and threshold, pass ratio and AUC before ICP; after ICP against the oracle's ICP
    (unpinned against open3d)."""
    from rcvpose_b200 import AccumulatorSpace as A
    g, tag = ycb_golden, "ycb_%d" % class_id
    root, name, n = _dataset(tmp_path, g, class_id)
    producer = lambda cls_name, k, image_path: synth.load_ycb_ckpt_maps(root, cls_name, k, image_path.split("/")[-2] + "_" + image_path.split("/")[-1].split("-")[0])  # noqa: E731
    res = A.estimate_6d_pose_ycb(types.SimpleNamespace(root_dataset=root, producer=producer, classes=[class_id], frames_per_batch=2))[name]
    assert res["frames"] == list(g[tag + "_frames"])
    assert np.array_equal(res["centre_mm"], g[tag + "_centres"])
    assert np.array_equal(res["n_points"], g[tag + "_n_points"])
    np.testing.assert_allclose(res["RT"], g[tag + "_RT"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(res["dist_before"], g[tag + "_dist_before"], rtol=1e-9)
    np.testing.assert_allclose(res["mean_before"], g[tag + "_icp_threshold"], rtol=1e-9)
    assert np.array_equal(res["scene_points"], g[tag + "_scene_points"])
    assert res["auc_before"] == pytest.approx(g[tag + "_summary"][0], abs=1e-12) and res["add_before"] == g[tag + "_summary"][2]
    if not sym:   # (ADD-S classes hand ICP a threshold of a fraction of a millimetre: one or two correspondences, no defined parity)
        np.testing.assert_allclose(res["dist_after"], g[tag + "_dist_after"], rtol=1e-5)
        assert res["auc_after"] == pytest.approx(g[tag + "_summary"][1], abs=1e-9) and res["add_after"] == g[tag + "_summary"][3]
    text = capsys.readouterr().out
    assert "ADD\\(s\\) AUC of " + name + " before ICP: " in text and "ADD\\(s\\) of " + name + " after ICP: " in text


def test_ycb_evaluator_needs_a_producer(tmp_path):
    root = str(tmp_path) + "/"
    synth.write_ycb_dataset(root, 5, evaluate.ycb_cls_names[5], 1, seed=2)
    with pytest.raises(ValueError):
        evaluate.evaluate_ycb_class(root, 5, None)
