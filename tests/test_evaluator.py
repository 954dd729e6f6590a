"""The rows either side of the voting path: N4 file formats, N1 the batched LINEMOD evaluator, N3 the
scene cloud + point-to-point ICP + ADD(-S) after it.  CPU tests cover the readers, the dataset layout and the oracle's ICP
restatement; the GPU tests compare the CUDA path (through the C ABI) with the reference's per-image loop restated on the
oracle's functions."""
import os
import struct
import types

import numpy as np
import pytest
import torch

from oracle import oracle
from rcvpose_b200 import formats, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------------
# N4 formats (CPU)
# ------------------------------------------------------------------------------------------------
def test_read_depth_dpt_roundtrip_and_errors(tmp_path):
    d = (np.arange(480 * 640, dtype=np.uint32) % 2000).astype(np.uint16).reshape(480, 640)
    p = str(tmp_path / "depth7.dpt")
    formats.write_depth_dpt(p, d)
    assert os.path.getsize(p) == 8 + 2 * d.size          # uint32 h, w + uint16 data (AccumulatorSpace.py:484-486)
    got = formats.read_depth(p)
    assert got.dtype == np.uint16 and got.shape == (480, 640) and np.array_equal(got, d)
    with open(p, "r+b") as f:
        f.truncate(8 + 100)
    with pytest.raises(ValueError):
        formats.read_depth(p)
    from PIL import Image
    q = str(tmp_path / "depth_00001.png")
    Image.fromarray(d).save(q)                            # the LMO / YCB branch: any other extension is an image (:488-489)
    assert np.array_equal(formats.read_depth(q), d)


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_read_ply_points_with_extra_properties_and_faces(tmp_path, fmt):
    rng = np.random.default_rng(1)
    n = 37
    xyz = rng.normal(size=(n, 3)).astype(np.float32)
    nrm = rng.normal(size=(n, 3)).astype(np.float32)
    rgb = rng.integers(0, 255, size=(n, 3)).astype(np.uint8)
    hdr = ("ply\nformat %s 1.0\ncomment made by a test\nelement vertex %d\nproperty float x\nproperty float y\nproperty float z\n"
           "property float nx\nproperty float ny\nproperty float nz\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
           "element face 2\nproperty list uchar int vertex_indices\nend_header\n" % (fmt, n))
    p = str(tmp_path / "m.ply")
    with open(p, "wb") as f:
        f.write(hdr.encode())
        if fmt == "ascii":
            for i in range(n):
                f.write(("%r %r %r %r %r %r %d %d %d\n" % (*map(float, xyz[i]), *map(float, nrm[i]), *rgb[i])).encode())
            f.write(b"3 0 1 2\n3 1 2 3\n")
        else:
            e = "<" if fmt == "binary_little_endian" else ">"
            for i in range(n):
                f.write(struct.pack(e + "6f3B", *xyz[i], *nrm[i], *rgb[i]))
            f.write(struct.pack(e + "B3i", 3, 0, 1, 2) + struct.pack(e + "B3i", 3, 1, 2, 3))
    got = formats.read_ply_points(p)
    assert got.dtype == np.float64 and got.shape == (n, 3)
    assert np.array_equal(got, xyz.astype(np.float64))


def test_read_ply_rejects_garbage(tmp_path):
    p = str(tmp_path / "x.ply")
    open(p, "wb").write(b"not a ply\n")
    with pytest.raises(ValueError):
        formats.read_ply_points(p)
    open(p, "wb").write(b"ply\nformat binary_little_endian 1.0\nelement vertex 5\nproperty float x\nproperty float y\nproperty float z\nend_header\n\0\0")
    with pytest.raises(ValueError):
        formats.read_ply_points(p)


def test_ycb_radius_h5_reader(tmp_path):
    """The reference's YCB generator stores radius maps in HDF5 (3DRadius_ycb.py:200-251); the reader needs h5py like the
    reference does.  With h5py: round trip of the generator's layout.  Without (this image): a clear ImportError."""
    p = str(tmp_path / "002_master_chef_can.hdf5")
    try:
        import h5py
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            formats.read_ycb_radius_h5(p, "0048_000001")
        return
    rng = np.random.default_rng(3)
    maps = rng.random((9, 12, 16)) * 10.0
    img = rng.integers(0, 255, (12, 16, 3), dtype=np.uint8)
    with h5py.File(p, "a") as f:
        f.create_group("/JPEGImages/").create_dataset("0048_000001", data=img, compression="gzip", compression_opts=9)
        for k in range(9):
            f.create_group("/3Dradius_pt%d_dm/" % k).create_dataset("0048_000001", data=maps[k], compression="gzip", compression_opts=9)
    got, im = formats.read_ycb_radius_h5(p, "0048_000001", with_image=True)
    assert got.dtype == np.float32 and got.shape == (3, 12, 16)
    assert np.array_equal(got, maps[1:4].astype(np.float32)) and np.array_equal(im, img)
    with pytest.raises(KeyError):
        formats.read_ycb_radius_h5(p, "0048_000002")


def test_linemod_layout_and_max_radii(tmp_path):
    from rcvpose_b200 import evaluate
    root = str(tmp_path) + "/"
    stems = synth.write_lm_dataset(root, "cat", 3, seed=5)
    cls = evaluate.LinemodClass(root, "cat")
    assert cls.stems == sorted(stems) and len(os.listdir(root + "LINEMOD/cat/JPEGImages")) == 5     # 2 images are not in val.txt
    assert cls.cad_m.shape == (1500, 3) and cls.keypoints_m.shape == (9, 3)
    # AccumulatorSpace.py:539-545
    want = [np.sqrt(((cls.cad_m - cls.keypoints_m[i + 1]) ** 2).sum(1)).max() * 10 for i in range(3)]
    np.testing.assert_allclose(cls.max_radii_dm, want, rtol=1e-15)
    d = cls.depth(stems[0])
    assert d.dtype == np.uint16 and d.shape == (480, 640) and (d != 0).sum() > 1000
    rt = cls.pose_mm(stems[0])
    raw = np.load(root + "LINEMOD/cat/pose/pose%d.npy" % int(stems[0]))
    assert np.array_equal(rt[:, :3], raw[:, :3]) and np.array_equal(rt[:, 3], raw[:, 3] * 1000)
    assert cls.radial_est(stems[0], 2).dtype == np.float32
    assert evaluate.add_threshold["ape"] == 0.01421240983190395 and evaluate.lm_syms == ["eggbox", "glue"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_evaluator_fails_loudly_without_gpu(tmp_path):
    from rcvpose_b200 import AccumulatorSpace as A
    root = str(tmp_path) + "/"
    synth.write_lm_dataset(root, "ape", 1)
    with pytest.raises(Exception) as ei:
        A.estimate_6d_pose_lm(types.SimpleNamespace(root_dataset=root, using_ckpts=False, classes=["ape"]))
    assert "CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


# ------------------------------------------------------------------------------------------------
# the oracle's restatements for these rows (CPU)
# ------------------------------------------------------------------------------------------------
def _pose(rv, t):
    th = np.linalg.norm(rv)
    k = rv / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    RT = np.eye(4)
    RT[:3, :3] = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    RT[:3, 3] = t
    return RT


def test_oracle_scene_union_matches_the_reference_loop():
    rng = np.random.default_rng(2)
    pool = rng.integers(0, 50, size=(40, 3)).astype(np.float64)
    clouds = [pool[rng.permutation(40)[:25]] for _ in range(3)]
    # the reference's loop, literally (AccumulatorSpace.py:620-625)
    ref = clouds[0]
    for c in clouds[1:]:
        for coor in c:
            if not (coor == ref).all(1).any():
                ref = np.append(ref, np.expand_dims(coor, axis=0), axis=0)
    assert np.array_equal(oracle.scene_union(clouds), ref)


def test_oracle_icp_recovers_a_rigid_motion_and_handles_no_correspondences():
    rng = np.random.default_rng(3)
    src = rng.normal(size=(800, 3)) * np.array([40.0, 25.0, 55.0])
    T = _pose(np.array([0.02, -0.03, 0.025]), np.array([1.5, -2.0, 1.0]))
    tgt = src @ T[:3, :3].T + T[:3, 3]
    reg = oracle.registration_icp(src, tgt, 15.0, np.eye(4), max_iteration=60)
    np.testing.assert_allclose(reg["transformation"], T, atol=1e-6)
    assert reg["fitness"] == 1.0 and reg["inlier_rmse"] < 1e-6
    far = oracle.registration_icp(src, tgt + 1e4, 1.0, np.eye(4))
    assert far["fitness"] == 0.0 and far["inlier_rmse"] == 0.0 and far["iterations"] == 1
    assert np.array_equal(far["transformation"], np.eye(4))
    # umeyama never returns a reflection
    a = rng.normal(size=(4, 3))
    b = a * np.array([1, 1, -1.0])
    assert np.linalg.det(oracle.umeyama_rigid(a, b)[:3, :3]) > 0


# ------------------------------------------------------------------------------------------------
# GPU: scene cloud, ICP, evaluator
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    from rcvpose_b200 import api
    return api.VoteContext(0, max_items=64, max_points_total=1 << 21, max_grid=400)


def _rows_sorted(a):
    a = np.ascontiguousarray(a)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


@pytest.mark.gpu
def test_scene_clouds_vs_reference_union(ctx):
    """xyz_mm_icp (AccumulatorSpace.py:612-625): same point SET as the reference's append-if-new loop, bit for bit, in
    row-major pixel order; ragged batch with an empty frame; capacity overflow is a status, not UB."""
    from rcvpose_b200 import api
    frames = [synth.config3_frame(f) for f in (21, 22, 23)]
    frames[1]["radius"][2, ::2] = 0                        # keypoint 3 sees half the rows: the union is not any single mask
    frames[2]["radius"][:] = 0                             # empty frame
    depth = np.stack([f["depth"] for f in frames])
    radius = np.stack([f["radius"] for f in frames])
    mr = np.stack([f["max_radii_dm"] for f in frames])
    K = frames[0]["K"]
    xyz, offs, st = ctx.scene_clouds(torch.from_numpy(depth.view(np.int16)).cuda(), torch.from_numpy(radius).cuda(), torch.from_numpy(K).cuda(),
                                     max_radii=torch.from_numpy(mr).cuda(), mask_flags=api.MASK_LM_NPY)
    xyz, offs, st = xyz.cpu().numpy(), offs.cpu().numpy(), st.cpu().numpy()
    assert offs[0] == 0 and list(st) == [api.RCV_ST_OK, api.RCV_ST_OK, api.RCV_ST_EMPTY_MASK] and offs[3] == offs[2]
    for b, fr in enumerate(frames[:2]):
        clouds = []
        for k in range(3):
            r = np.where(radius[b, k] <= mr[b, k], radius[b, k], 0)
            clouds.append(oracle.rgbd_to_point_cloud(K, depth[b] * np.where(r != 0, 1, 0)))
        want = oracle.scene_union(clouds)
        got = xyz[offs[b]:offs[b + 1]]
        assert got.shape == want.shape and len(want) > max(len(c) for c in clouds) - 1
        assert np.array_equal(_rows_sorted(got), _rows_sorted(want))
        # row-major order = rgbd_to_point_cloud of the OR of the masks
        m = np.zeros(depth[b].shape, bool)
        for k in range(3):
            m |= (radius[b, k] <= mr[b, k]) & (radius[b, k] != 0)
        assert np.array_equal(got, oracle.rgbd_to_point_cloud(K, depth[b] * m))
    # scale (the YCB evaluator's xyz_icp*1000, :1154) and a capacity that holds only the first frame
    n0 = int(offs[1])
    xyz2, offs2, st2 = ctx.scene_clouds(torch.from_numpy(depth.view(np.int16)).cuda(), torch.from_numpy(radius).cuda(), torch.from_numpy(K).cuda(),
                                        max_radii=torch.from_numpy(mr).cuda(), mask_flags=api.MASK_LM_NPY, scale=1000.0, capacity=n0 + 5)
    assert np.array_equal(xyz2[:n0].cpu().numpy(), xyz[:n0] * 1000.0)
    assert list(offs2.cpu().numpy()) == [0, n0, n0, n0] and int(st2[1]) == api.RCV_ST_POINT_OVERFLOW


@pytest.mark.gpu
def test_icp_vs_oracle(ctx):
    """registration_icp, point to point (AccumulatorSpace.py:704-718): the CUDA loop against the oracle's restatement of open3d's
    algorithm (k-d tree + SVD umeyama) on ragged scenes: same iteration counts and fitness, transformation / rmse to 1e-7
    (nearest neighbours are exact on both sides; the rotation comes from Horn's quaternion instead of an SVD)."""
    rng = np.random.default_rng(41)
    M = 1100
    u = rng.normal(size=(M, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    model = u * np.array([45.0, 30.0, 60.0])
    cases = []
    for b in range(7):
        gt = _pose(rng.normal(size=3), rng.uniform(-100, 100, 3) + np.array([0, 0, 900.0]))
        pts = model @ gt[:3, :3].T + gt[:3, 3]
        vis = pts[pts[:, 2] < np.median(pts[:, 2])]                                   # the camera-facing half
        scene = vis[rng.permutation(len(vis))[: 200 + 57 * b]] + rng.normal(0, 0.3, size=(200 + 57 * b, 3))
        init = _pose(rng.normal(size=3) * 0.02, rng.normal(size=3) * 1.5) @ gt
        cases.append((scene, init, [6.0, 3.0, 10.0, 4.0, 8.0, 1e-3, 5.0][b]))
    cases[6] = (np.zeros((0, 3)), cases[6][1], 5.0)                                    # empty scene: nothing to register
    offs = np.cumsum([0] + [len(c[0]) for c in cases])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    for max_iter in (30, 100):
        reg = ctx.icp(t(model), t(np.concatenate([c[0] for c in cases])), t(offs), t(np.stack([c[1] for c in cases])),
                      t(np.array([c[2] for c in cases])), max_iter=max_iter)
        reg = {k: v.cpu().numpy() for k, v in reg.items()}
        for b, (scene, init, md) in enumerate(cases):
            want = oracle.registration_icp(model, scene, md, init, max_iteration=max_iter)
            assert reg["iters"][b] == want["iterations"], (max_iter, b, reg["iters"][b], want["iterations"])
            assert abs(reg["fitness"][b] - want["fitness"]) < 1e-12, (b, reg["fitness"][b], want["fitness"])
            np.testing.assert_allclose(reg["rmse"][b], want["inlier_rmse"], rtol=1e-7, atol=1e-9)
            np.testing.assert_allclose(reg["RT"][b], want["transformation"], rtol=0, atol=1e-6)
            assert abs(np.linalg.det(reg["RT"][b][:3, :3]) - 1.0) < 1e-12 and np.array_equal(reg["RT"][b][3], [0, 0, 0, 1])
        assert reg["fitness"][5] < 0.05 and reg["fitness"][6] == 0.0 and np.allclose(reg["RT"][6], cases[6][1], atol=1e-15)
        assert 0.0 < reg["rmse"][0] < 6.0 and reg["fitness"][2] > 0.3
    # max_iter 0: open3d evaluates the initial pose and returns it
    reg0 = ctx.icp(t(model), t(np.concatenate([c[0] for c in cases])), t(offs), t(np.stack([c[1] for c in cases])),
                   t(np.array([c[2] for c in cases])), max_iter=0)
    assert int(reg0["iters"].max()) == 0 and np.allclose(reg0["RT"].cpu().numpy(), np.stack([c[1] for c in cases]), atol=0)


@pytest.mark.gpu
def test_icp_grid_search_equals_all_pairs(ctx, monkeypatch):
    """The uniform-grid correspondence search (refine.cu, default) against the all-pairs search (RCV_ICP_BRUTE=1): same scores,
    same tie rule, every scene point within max_dist inside the visited cells -- bit-identical transformations, fitness, rmse
    and iteration counts, for thresholds from far below the point spacing to larger than the scene."""
    rng = np.random.default_rng(43)
    M = 3000
    u = rng.normal(size=(M, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    model = u * np.array([60.0, 35.0, 80.0])
    scenes, inits, mds = [], [], []
    for b, md in enumerate([0.05, 1.0, 4.0, 9.0, 25.0, 120.0, 1.0e4, 6.0, 6.0]):
        gt = _pose(rng.normal(size=3), rng.uniform(-100, 100, 3) + np.array([0, 0, 900.0]))
        pts = model @ gt[:3, :3].T + gt[:3, 3]
        vis = pts[pts[:, 2] < np.median(pts[:, 2])]
        n = 700 + 130 * b
        sc = vis[rng.integers(0, len(vis), size=n)] + rng.normal(0, 0.4, size=(n, 3))
        if b == 7:
            sc = np.repeat(sc[:300], 3, axis=0)                      # exact duplicates: ties between scene indices
        if b == 8:
            sc = sc[:1]                                              # a single scene point: a one-cell grid
        scenes.append(sc); mds.append(md)
        inits.append(_pose(rng.normal(size=3) * 0.03, rng.normal(size=3) * 2.0) @ gt)
    offs = np.cumsum([0] + [len(x) for x in scenes])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    args = (t(model), t(np.concatenate(scenes)), t(offs), t(np.stack(inits)), t(np.array(mds)))
    monkeypatch.setenv("RCV_ICP_BRUTE", "1")
    want = {k: v.cpu().numpy() for k, v in ctx.icp(*args, max_iter=30).items()}
    monkeypatch.setenv("RCV_ICP_BRUTE", "0")
    got = {k: v.cpu().numpy() for k, v in ctx.icp(*args, max_iter=30).items()}
    for k in ("RT", "fitness", "rmse", "iters"):
        assert np.array_equal(got[k], want[k]), k
    assert got["fitness"][3] > 0.2 and got["fitness"][0] < got["fitness"][3]


def _reference_loop(root, class_name, sym, threshold_mm):
    """estimate_6d_pose_lm's per-image loop (AccumulatorSpace.py:553-731, npy branch) on the oracle's functions."""
    from rcvpose_b200 import evaluate
    cls = evaluate.LinemodClass(root, class_name)
    K = evaluate.linemod_K
    out = []
    for stem in cls.stems:
        depth1 = cls.depth(stem)
        est = np.zeros((3, 3))
        icp_clouds = []
        for k in (1, 2, 3):
            r = cls.radial_est(stem, k)
            r = np.where(r <= cls.max_radii_dm[k - 1], r, 0)
            dm = depth1 * np.where(r != 0, 1, 0)
            xyz_mm = oracle.rgbd_to_point_cloud(K, dm)
            est[k - 1] = oracle.Accumulator_3D(xyz_mm / 1000, r[dm.nonzero()])[0]
            icp_clouds.append(xyz_mm)
        RT = np.zeros((4, 4))
        oracle.lmshorn(cls.keypoints_m[1:4] * 1000, est, 3, RT)
        gt = np.eye(4)
        gt[:3] = cls.pose_mm(stem)
        mean, mn = oracle.add_metric(cls.cad_m * 1000, RT, gt)
        before = mn if sym else mean
        reg = oracle.registration_icp(cls.cad_m * 1000, oracle.scene_union(icp_clouds), before, RT)
        mean2, mn2 = oracle.add_metric(cls.cad_m * 1000, reg["transformation"], gt)
        after = mn2 if sym else mean2
        out.append(dict(centres=est, RT=RT, before=before, after=after, RT_icp=reg["transformation"], iters=reg["iterations"],
                        fitness=reg["fitness"], pb=before <= threshold_mm, pa=after <= threshold_mm))
    return cls, out


@pytest.mark.gpu
@pytest.mark.parametrize("class_name,dtype", [("ape", np.float32), ("eggbox", np.float32), ("cat", np.float64)])
def test_estimate_6d_pose_lm_vs_reference_loop(tmp_path, class_name, dtype, capsys):
    """The drop-in evaluator on a synthetic class laid out like LINEMOD (a non-symmetric class: mean distance; a symmetric one:
    minimum distance, :688-695; float64 radius maps keep their dtype through the exact drop-in surface) against the reference's
    per-image loop: keypoints bit-identical, Horn pose 1e-9, ADD(-S) before ICP 1e-9 relative, ICP iteration count equal,
    refined pose 1e-6, ADD(-S) after ICP 1e-6 relative, pass / fail counters equal (ICP parity for the non-symmetric classes)."""
    from rcvpose_b200 import AccumulatorSpace as A, evaluate
    root = str(tmp_path) + "/"
    synth.write_lm_dataset(root, class_name, 3, seed=len(class_name))
    opts = types.SimpleNamespace(root_dataset=root, using_ckpts=False, classes=[class_name], frames_per_batch=2)
    res = A.estimate_6d_pose_lm(opts)[class_name]
    sym = class_name in evaluate.lm_syms
    cls, want = _reference_loop(root, class_name, sym, evaluate.add_threshold[class_name] * 1000)
    assert res["frames"] == cls.stems and res["n"] == 3
    for i, w in enumerate(want):
        assert np.array_equal(res["centre_mm"][i], w["centres"]), (i, res["centre_mm"][i], w["centres"])
        np.testing.assert_allclose(res["RT"][i], w["RT"], rtol=0, atol=1e-9)
        assert abs(res["dist_before"][i] - w["before"]) <= 1e-9 * max(1.0, w["before"])
        assert bool(res["passed_before"][i]) == bool(w["pb"])
        R = res["RT_icp"][i][:3, :3]
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12 and np.isfinite(res["dist_after"][i])
        if sym:
            # The reference's threshold for a symmetric class is the MINIMUM nearest-neighbour distance (:706-707), a fraction of a
            # millimetre: ICP then sees one or two correspondences, the rotation is not determined by the data, and open3d's own
            # answer depends on the null-space basis its SVD happens to return -- no parity is defined for the refined pose.
            assert res["icp_fitness"][i] * len(cls.cad_m) < 50
            continue
        assert res["icp_iters"][i] == w["iters"] and abs(res["icp_fitness"][i] - w["fitness"]) < 1e-12
        np.testing.assert_allclose(res["RT_icp"][i], w["RT_icp"], rtol=0, atol=1e-6)
        assert abs(res["dist_after"][i] - w["after"]) <= 1e-6 * max(1.0, w["after"])
        assert bool(res["passed_after"][i]) == bool(w["pa"])
    assert res["add_before"] == sum(w["pb"] for w in want) / 3
    assert sym or res["add_after"] == sum(w["pa"] for w in want) / 3
    text = capsys.readouterr().out
    assert ("ADDs of " if sym else "ADD of ") + class_name + " before ICP: " in text and "Evaluation on  " + class_name in text
    # the synthetic radii are accurate to ~1 mm: the keypoints land within a voxel or two of the truth and the pose passes ADD
    assert res["passed_before"].all() or sym


@pytest.mark.gpu
def test_evaluator_empty_mask_raises_like_the_reference(tmp_path):
    from rcvpose_b200 import AccumulatorSpace as A
    root = str(tmp_path) + "/"
    stems = synth.write_lm_dataset(root, "duck", 2, seed=9)
    p = root + "LINEMOD_ORIG/estRadialMap/duck/Out_pt2_dm/" + stems[1] + ".npy"
    np.save(p, np.zeros((480, 640), np.float32))
    with pytest.raises(ValueError) as ei:       # Accumulator_3D on an empty cloud: ValueError from .min()
        A.estimate_6d_pose_lm(types.SimpleNamespace(root_dataset=root, using_ckpts=False, classes=["duck"]))
    assert "zero-size array" in str(ei.value) and stems[1] in str(ei.value)


@pytest.mark.gpu
def test_fused_head_vote_frames_bit_identical_to_two_calls(ctx):
    """conv8 -> mask rule -> voting in one entry point (the head's epilogue writes the survival bits; the seg plane
    never reaches HBM) against rcv_head_1x1 followed by rcv_vote_frames on the maps it produced: every output bit-identical,
    the survivors equal to the mask rule evaluated in NumPy on those maps, and one item checked against the oracle."""
    from rcvpose_b200 import api
    frames = [synth.config3_frame(f) for f in (31, 32, 33)]
    B, Kp, H, W = 3, 3, 480, 640
    g = torch.Generator(device="cuda").manual_seed(7)
    up = torch.relu(torch.randn((B, Kp, 32, H, W), generator=g, device="cuda")) * 0.5
    depth_np = np.stack([f["depth"] for f in frames])
    rad_np = np.stack([f["radius"] for f in frames])
    up[:, :, 0] = torch.from_numpy(rad_np).cuda()                                                # channel 0 carries the radius (dm)
    inside = torch.from_numpy((depth_np != 0).astype(np.float32)).cuda()[:, None].expand(B, Kp, H, W)
    up[:, :, 1] = inside * (0.75 + 0.2 * torch.rand((B, Kp, H, W), generator=g, device="cuda"))   # channel 1: seg score around the 0.8 threshold
    up = up.to(torch.bfloat16).contiguous()
    weight = torch.randn((Kp, 2, 32), generator=g, device="cuda") * 2e-4
    weight[:, 0, 1] = 1.0
    weight[:, 1, 0] = 1.0
    bias = torch.randn((Kp, 2), generator=g, device="cuda") * 1e-3
    depth = torch.from_numpy(depth_np.view(np.int16)).cuda()
    K = torch.from_numpy(frames[0]["K"]).cuda()
    mr = torch.from_numpy(np.stack([f["max_radii_dm"] for f in frames])).cuda()
    fused = ctx.head_vote_frames(up, weight, bias, depth, K, max_radii=mr, mask_flags=api.MASK_LM_CKPT, want_radius=True)
    maps = torch.stack([ctx.head_1x1(up[:, k].contiguous(), weight[k], bias[k]) for k in range(Kp)], dim=1)    # (B,Kp,2,H,W)
    sem, radius = maps[:, :, 0].contiguous(), maps[:, :, 1].contiguous()
    two = ctx.vote_frames(depth, radius, K, sem=sem, max_radii=mr, mask_flags=api.MASK_LM_CKPT)
    torch.cuda.synchronize()
    assert torch.equal(fused["radius"], radius)
    for k in ("centre_mm", "peak", "votes", "n_points", "grid", "status"):
        assert torch.equal(fused[k], two[k]), k
    assert int(fused["status"].abs().sum()) == 0 and int(fused["votes"].min()) > 0
    sem_np, radius_np = sem.cpu().numpy(), radius.cpu().numpy()
    m = (depth_np[:, None] != 0) & (sem_np > np.float32(0.8)) & (radius_np.astype(np.float64) <= mr.cpu().numpy()[:, :, None, None])
    assert np.array_equal(fused["n_points"].cpu().numpy(), m.sum(axis=(2, 3)))
    assert 0.2 < m[0, 0].sum() / (depth_np[0] != 0).sum() < 0.8            # the seg threshold really cuts
    xyz_mm = oracle.rgbd_to_point_cloud(frames[1]["K"], depth_np[1] * m[1, 2])
    want = oracle.Accumulator_3D(xyz_mm / 1000, radius_np[1, 2][(depth_np[1] * m[1, 2]).nonzero()])
    assert np.array_equal(fused["centre_mm"][1, 2].cpu().numpy(), want[0])
    # a second context-scratch call (no radius_out) gives the same results
    again = ctx.head_vote_frames(up, weight, bias, depth, K, max_radii=mr, mask_flags=api.MASK_LM_CKPT)
    assert torch.equal(again["centre_mm"], fused["centre_mm"]) and torch.equal(again["votes"], fused["votes"])


def test_lmo_layout_counts_every_directory_entry(tmp_path):
    from rcvpose_b200 import evaluate
    root = str(tmp_path) + "/"
    stems = synth.write_lmo_dataset(root, "duck", 3, seed=4)
    cls = evaluate.LmoClass(root, "duck")
    assert cls.stems == stems and len(cls.entries) == 5          # 3 complete + no pose + no third map (:962 counts all)
    assert cls.index("color_00076") == 76
    d = cls.depth(stems[0])
    assert d.dtype == np.float64 and d.shape == (480, 640) and (d != 0).sum() > 1000     # :833
    assert len(evaluate.LmoClass(root, "duck", using_ckpts=True).stems) == 4              # the ckpt branch only needs the pose (:809-811)


def _reference_loop_lmo(root, class_name, sym, thr):
    """estimate_6d_pose_lmo's per-image loop (AccumulatorSpace.py:786-962, npy branch) on the oracle's functions."""
    from rcvpose_b200 import evaluate
    cls = evaluate.LmoClass(root, class_name)
    K = evaluate.linemod_K
    out = []
    for stem in cls.stems:
        depth_map1 = cls.depth(stem)
        est = np.zeros((3, 3))
        clouds = []
        for k in (1, 2, 3):
            r = np.load(cls.radial_path(stem, k))
            r = np.where(r <= cls.max_radii_dm[k - 1], r, 0)
            sem = np.where(r > 0, 1, 0)
            dm = depth_map1 * sem
            if r.max() != 0:
                xyz_mm = oracle.rgbd_to_point_cloud(K, dm)
                est[k - 1] = oracle.Accumulator_3D(xyz_mm / 1000, r[dm.nonzero()])[0]
                clouds.append(xyz_mm)
        RT = np.zeros((4, 4))
        oracle.lmshorn(cls.keypoints_m[1:4] * 1000, est, 3, RT)
        gt = np.eye(4)
        gt[:3] = cls.pose_mm(stem)
        mean, mn = oracle.add_metric(cls.cad_m * 1000, RT, gt)
        before = mn if sym else mean
        reg = oracle.registration_icp(cls.cad_m * 1000, oracle.scene_union(clouds), before, RT, relative_fitness=thr, relative_rmse=thr)
        mean2, mn2 = oracle.add_metric(cls.cad_m * 1000, reg["transformation"], gt)
        after = mn2 if sym else mean2
        out.append(dict(centres=est, RT=RT, before=before, after=after, RT_icp=reg["transformation"], iters=reg["iterations"],
                        pb=before <= thr, pa=after <= thr))
    return cls, out


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_estimate_6d_pose_lmo_vs_reference_loop(tmp_path, dtype):
    """The Occlusion-LINEMOD drop-in (float64 depth images, `radial > 0` rule, skipped all-zero keypoint with its row left at 0,
    ICP criteria add_threshold*1000, ratios over every directory entry) against the reference's per-image loop on the oracle."""
    from rcvpose_b200 import AccumulatorSpace as A, evaluate
    root = str(tmp_path) + "/"
    synth.write_lmo_dataset(root, "can", 3, seed=6, radius_dtype=dtype)
    res = A.estimate_6d_pose_lmo(types.SimpleNamespace(root_dataset=root, using_ckpts=False, classes=["can"], frames_per_batch=2))["can"]
    thr = evaluate.add_threshold["can"] * 1000
    cls, want = _reference_loop_lmo(root, "can", False, thr)
    assert res["frames"] == cls.stems and res["n"] == 5 and res["evaluated"] == 3
    assert np.array_equal(res["centre_mm"][1, 1], np.zeros(3)) and np.array_equal(want[1]["centres"][1], np.zeros(3))
    for i, w in enumerate(want):
        assert np.array_equal(res["centre_mm"][i], w["centres"]), (i, res["centre_mm"][i], w["centres"])
        np.testing.assert_allclose(res["RT"][i], w["RT"], rtol=0, atol=1e-9)
        assert abs(res["dist_before"][i] - w["before"]) <= 1e-9 * max(1.0, w["before"])
        assert res["icp_iters"][i] == w["iters"] <= 2                  # criteria of ~28 mm: the first or second update already "converges"
        np.testing.assert_allclose(res["RT_icp"][i], w["RT_icp"], rtol=0, atol=1e-6)
        assert abs(res["dist_after"][i] - w["after"]) <= 1e-6 * max(1.0, w["after"])
        assert bool(res["passed_before"][i]) == bool(w["pb"]) and bool(res["passed_after"][i]) == bool(w["pa"])
    assert res["add_before"] == sum(w["pb"] for w in want) / 5 and res["add_after"] == sum(w["pa"] for w in want) / 5
    assert res["passed_before"][0] and not res["passed_before"][1]     # a keypoint at the origin ruins the pose of frame 1


def _gloo_eval_worker(rank, world, port, q):
    import torch.distributed as dist
    from rcvpose_b200 import evaluate
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stems = ["%06d" % i for i in range(7)]
        mine = evaluate.shard_frames(stems)
        counts = evaluate.reduce_counts([len(mine), rank + 1, 10 * rank])
        q.put((rank, mine, counts))
    finally:
        dist.destroy_process_group()


def test_sharded_evaluation_world2_gloo():
    """N > 1 evaluation on CPU: two ranks take contiguous frame ranges that partition the class, and the pass counters are
    summed with one small all_reduce (the evaluators' only communication)."""
    import socket
    import torch.multiprocessing as mp
    from rcvpose_b200 import evaluate
    assert evaluate.shard_frames(["a", "b", "c"]) == ["a", "b", "c"] and evaluate.reduce_counts([3, 4]) == [3, 4]     # single process
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] + res[1][1] == ["%06d" % i for i in range(7)] and abs(len(res[0][1]) - len(res[1][1])) == 1
    assert res[0][2] == res[1][2] == [7, 3, 10]


def test_evaluator_needs_a_producer_for_the_checkpoint_branch(tmp_path):
    from rcvpose_b200 import evaluate
    root = str(tmp_path) + "/"
    synth.write_lm_dataset(root, "ape", 1)
    with pytest.raises(ValueError) as ei:       # checked before any GPU work: the network is not part of the package
        evaluate.evaluate_lm_class(root, "ape", using_ckpts=True)
    assert "producer" in str(ei.value)


# ------------------------------------------------------------------------------------------------
# pinned against the REAL reference's evaluators (tests/golden/make_golden_evaluator.py ran estimate_6d_pose_lm / _lmo of
# aaronWool/rcvpose, unmodified, in the build container on these same synthetic datasets; open3d replaced by a stand-in)
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eval_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "evaluator_golden.npz"))


def _check_lm_against_golden(g, cls, got, sym, icp_tol=1e-6, check_icp=True):
    """got: list of dicts(centres, RT, before, after, RT_icp, iters) in sorted-stem order."""
    tag = "lm_" + cls
    assert len(got) == len(g[tag + "_stems"])
    for i, w in enumerate(got):
        assert np.array_equal(w["centres"], g[tag + "_centres"][i]), (cls, i)                  # the reference's Accumulator_3D outputs
        np.testing.assert_allclose(w["RT"], g[tag + "_RT"][i], rtol=0, atol=1e-9)             # the reference's lmshorn output
        assert abs(w["before"] - g[tag + "_dist_before"][i]) <= 1e-9 * max(1.0, w["before"])  # project() + nearest neighbour, as the reference ran it
        if check_icp and not sym:    # the stand-in's ICP (the oracle's restatement) ran inside the reference's own loop
            assert w["iters"] == g[tag + "_icp_iters"][i]
            np.testing.assert_allclose(w["RT_icp"], g[tag + "_icp_RT"][i], rtol=0, atol=icp_tol)
            assert abs(w["after"] - g[tag + "_dist_after"][i]) <= icp_tol * max(1.0, w["after"])


@pytest.mark.parametrize("cls", ["ape", "eggbox"])
def test_oracle_loop_matches_the_real_reference_evaluator_lm(tmp_path, eval_golden, cls):
    """CPU: the per-image loop restated on the oracle (the checker of the GPU tests) against what the reference's own
    estimate_6d_pose_lm computed: keypoints bit-identical, poses 1e-9, ADD(-S) distances, ICP iteration counts, final ratios."""
    from rcvpose_b200 import evaluate
    n_frames, seed = (int(v) for v in eval_golden["lm_" + cls + "_seed"])
    root = str(tmp_path) + "/"
    stems = synth.write_lm_dataset(root, cls, n_frames, seed=seed)
    assert sorted(stems) == list(eval_golden["lm_" + cls + "_stems"])
    sym = cls in evaluate.lm_syms
    c, want = _reference_loop(root, cls, sym, evaluate.add_threshold[cls] * 1000)
    _check_lm_against_golden(eval_golden, cls, want, sym, icp_tol=1e-9)
    assert np.array_equal(eval_golden["lm_" + cls + "_scene_points"], [len(oracle.scene_union(
        [oracle.rgbd_to_point_cloud(evaluate.linemod_K, c.depth(s) * np.where(np.where(c.radial_est(s, k) <= c.max_radii_dm[k - 1], c.radial_est(s, k), 0) != 0, 1, 0))
         for k in (1, 2, 3)])) for s in c.stems])                                              # size of the reference's xyz_mm_icp
    rb, ra = eval_golden["lm_" + cls + "_ratios"]
    assert rb == sum(w["pb"] for w in want) / n_frames and (sym or ra == sum(w["pa"] for w in want) / n_frames)


def _match_lmo(g, rows):
    """The reference walks os.listdir order: pair each of its frames with the row that has the same estimated keypoints."""
    pairs = []
    for j in range(len(g["lmo_can_est_kpts"])):
        hit = [i for i, w in enumerate(rows) if np.array_equal(w["centres"], g["lmo_can_est_kpts"][j])]
        assert len(hit) == 1, (j, hit)
        pairs.append((hit[0], j))
    assert sorted(i for i, _ in pairs) == list(range(len(rows)))
    return pairs


def test_oracle_loop_matches_the_real_reference_evaluator_lmo(tmp_path, eval_golden):
    from rcvpose_b200 import evaluate
    n_frames, seed = (int(v) for v in eval_golden["lmo_can_seed"])
    root = str(tmp_path) + "/"
    synth.write_lmo_dataset(root, "can", n_frames, seed=seed)
    thr = evaluate.add_threshold["can"] * 1000
    cls, want = _reference_loop_lmo(root, "can", False, thr)
    for i, j in _match_lmo(eval_golden, want):
        np.testing.assert_allclose(want[i]["RT"], eval_golden["lmo_can_RT"][j], rtol=0, atol=1e-9)
        assert want[i]["iters"] == eval_golden["lmo_can_icp_iters"][j]
        np.testing.assert_allclose(want[i]["RT_icp"], eval_golden["lmo_can_icp_RT"][j], rtol=0, atol=1e-9)
    rb, ra = eval_golden["lmo_can_ratios"]
    assert rb == sum(w["pb"] for w in want) / len(cls.entries) and ra == sum(w["pa"] for w in want) / len(cls.entries)


_pending = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def _pending_gpu(fn):
    for m in _pending:
        fn = m(fn)
    return fn


@_pending_gpu
@pytest.mark.parametrize("cls", ["ape", "eggbox"])
def test_estimate_6d_pose_lm_vs_real_reference_golden(tmp_path, eval_golden, cls):
    """GPU: the drop-in evaluator against the outputs of the reference's own estimate_6d_pose_lm on the same dataset -- the pinned
    part: keypoints bit-identical, Horn pose, ADD(-S) distance before ICP, size of the ICP target, pass ratios.  (ICP itself
    ran through a stand-in there; the CUDA ICP is compared with the oracle in test_icp_vs_oracle and the tests above.)"""
    from rcvpose_b200 import AccumulatorSpace as A, evaluate
    n_frames, seed = (int(v) for v in eval_golden["lm_" + cls + "_seed"])
    root = str(tmp_path) + "/"
    synth.write_lm_dataset(root, cls, n_frames, seed=seed)
    res = A.estimate_6d_pose_lm(types.SimpleNamespace(root_dataset=root, using_ckpts=False, classes=[cls], frames_per_batch=2))[cls]
    sym = cls in evaluate.lm_syms
    got = [dict(centres=res["centre_mm"][i], RT=res["RT"][i], before=res["dist_before"][i], after=res["dist_after"][i], RT_icp=res["RT_icp"][i],
                iters=res["icp_iters"][i]) for i in range(n_frames)]
    assert res["frames"] == list(eval_golden["lm_" + cls + "_stems"])
    _check_lm_against_golden(eval_golden, cls, got, sym, check_icp=False)
    assert np.array_equal(res["scene_points"], eval_golden["lm_" + cls + "_scene_points"])
    rb, ra = eval_golden["lm_" + cls + "_ratios"]
    assert res["add_before"] == rb and (sym or res["add_after"] == ra)


@_pending_gpu
def test_estimate_6d_pose_lmo_vs_real_reference_golden(tmp_path, eval_golden):
    from rcvpose_b200 import AccumulatorSpace as A
    n_frames, seed = (int(v) for v in eval_golden["lmo_can_seed"])
    root = str(tmp_path) + "/"
    synth.write_lmo_dataset(root, "can", n_frames, seed=seed)
    res = A.estimate_6d_pose_lmo(types.SimpleNamespace(root_dataset=root, using_ckpts=False, classes=["can"], frames_per_batch=2))["can"]
    rows = [dict(centres=res["centre_mm"][i]) for i in range(n_frames)]
    for i, j in _match_lmo(eval_golden, rows):
        np.testing.assert_allclose(res["RT"][i], eval_golden["lmo_can_RT"][j], rtol=0, atol=1e-9)
        assert res["scene_points"][i] == eval_golden["lmo_can_scene_points"][j]
    rb, ra = eval_golden["lmo_can_ratios"]
    assert res["add_before"] == rb and res["add_after"] == ra


# ------------------------------------------------------------------------------------------------
# host logic of the evaluators with the device stage chain replaced by a recorder (CPU)
# ------------------------------------------------------------------------------------------------
class _FakeEvaluator:
    calls = []

    def __init__(self, cad_mm, kpts_mm, symmetric, threshold_mm, **kw):
        self.cad_mm, self.kpts_mm, self.symmetric, self.threshold_mm, self.kw = cad_mm, kpts_mm, symmetric, threshold_mm, kw

    def run(self, depth, radius, K, gt, **kw):
        from rcvpose_b200 import api
        B = radius.shape[0]
        _FakeEvaluator.calls.append(dict(depth=depth, radius=radius, K=K, gt=gt, ev=self, **kw))
        status = np.zeros((B, 3), np.int32)
        if getattr(_FakeEvaluator, "empty_at", None) is not None and len(_FakeEvaluator.calls) == 1:
            status[_FakeEvaluator.empty_at] = api.RCV_ST_EMPTY_MASK
        before = np.arange(B, dtype=np.float64) + 10.0 * len(_FakeEvaluator.calls)
        return dict(centre_mm=np.zeros((B, 3, 3)), RT=np.tile(np.eye(4), (B, 1, 1)), dist_before=before, passed_before=before < 21.0,
                    status=status, n_points=np.ones((B, 3), np.int32), peak=np.ones((B, 3), np.int32), grid=np.ones((B, 3), np.int32),
                    RT_icp=np.tile(np.eye(4), (B, 1, 1)), dist_after=before / 2, passed_after=before / 2 < 21.0,
                    icp_fitness=np.ones(B), icp_rmse=np.zeros(B), icp_iters=np.ones(B, np.int32), scene_points=np.ones(B, np.int64))


def test_lm_evaluator_host_logic_checkpoint_branch_and_batching(tmp_path, monkeypatch, capsys):
    """The checkpoint branch (AccumulatorSpace.py:594-610) and the batching, with the device chain replaced by a recorder:
    the producer is asked for every (frame, keypoint) in order, its maps reach the chain with the sem > 0.8 + max-radius rule,
    poses are converted to mm, batches respect frames_per_batch, counters and the reference's print-out add up."""
    from rcvpose_b200 import api, evaluate
    root = str(tmp_path) + "/"
    stems = synth.write_lm_dataset(root, "lamp", 5, seed=2)
    monkeypatch.setattr(evaluate, "FrameEvaluator", _FakeEvaluator)
    _FakeEvaluator.calls, _FakeEvaluator.empty_at = [], None
    asked = []

    def producer(cls, k, path):
        asked.append((cls, k, os.path.basename(path)))
        return np.full((480, 640), 0.9, np.float32), np.full((480, 640), float(k), np.float32)
    res = evaluate.evaluate_lm_class(root, "lamp", using_ckpts=True, producer=producer, frames_per_batch=2)
    assert asked == [("lamp", k, s + ".jpg") for s in sorted(stems) for k in (1, 2, 3)]
    assert [c["radius"].shape[0] for c in _FakeEvaluator.calls] == [2, 2, 1]
    c0 = _FakeEvaluator.calls[0]
    assert c0["mask_flags"] == api.MASK_LM_CKPT and c0["sem"].shape == (2, 3, 480, 640) and float(c0["sem"][0, 0, 0, 0]) == np.float32(0.9)
    assert np.array_equal(c0["radius"][1, 2], np.full((480, 640), 3.0, np.float32)) and c0["depth"].dtype == np.uint16
    raw = np.load(root + "LINEMOD/lamp/pose/pose%d.npy" % int(sorted(stems)[0]))
    assert np.array_equal(c0["gt"][0][:, 3], raw[:, 3] * 1000) and np.array_equal(c0["gt"][0][:, :3], raw[:, :3])
    ev = c0["ev"]
    assert ev.symmetric is False and ev.threshold_mm == evaluate.add_threshold["lamp"] * 1000
    assert np.allclose(ev.kpts_mm, np.load(root + "LINEMOD/lamp/Outside9.npy")[1:4] * 1000) and ev.cad_mm.shape == (1500, 3)
    # before = [10, 11 | 20, 21 | 30]: 3 of 5 pass; after = before / 2: all pass
    assert res["n"] == 5 and res["add_before"] == 3 / 5 and res["add_after"] == 1.0 and res["frames"] == sorted(stems)
    out = capsys.readouterr().out
    assert "ADD of lamp before ICP:  0.6" in out and "ADD of lamp after ICP:  1.0" in out


def test_lm_evaluator_host_logic_empty_mask_is_the_reference_error(tmp_path, monkeypatch):
    from rcvpose_b200 import evaluate
    root = str(tmp_path) + "/"
    stems = synth.write_lm_dataset(root, "iron", 2, seed=3)
    monkeypatch.setattr(evaluate, "FrameEvaluator", _FakeEvaluator)
    _FakeEvaluator.calls, _FakeEvaluator.empty_at = [], (1, 2)
    with pytest.raises(ValueError) as ei:
        evaluate.evaluate_lm_class(root, "iron", frames_per_batch=2, verbose=False)
    assert sorted(stems)[1] in str(ei.value) and "keypoint 3" in str(ei.value)
    _FakeEvaluator.empty_at = None


def test_lmo_evaluator_host_logic_zero_maps_and_counters(tmp_path, monkeypatch):
    """LMO host logic with the recorder: all-zero thresholded maps are tolerated (the keypoint is skipped, :858), an empty cloud
    under surviving radii is the reference's ValueError, ICP criteria are add_threshold * 1000 (:939-941), ratios are over every
    directory entry (:962)."""
    from rcvpose_b200 import api, evaluate
    root = str(tmp_path) + "/"
    stems = synth.write_lmo_dataset(root, "glue", 3, seed=8)
    monkeypatch.setattr(evaluate, "FrameEvaluator", _FakeEvaluator)
    _FakeEvaluator.calls, _FakeEvaluator.empty_at = [], (1, 1)            # frame 1 / keypoint 2 has the all-zero map: tolerated
    res = evaluate.evaluate_lmo_class(root, "glue", frames_per_batch=8, verbose=False)
    c0 = _FakeEvaluator.calls[0]
    thr = evaluate.add_threshold["glue"] * 1000
    assert c0["mask_flags"] == api.MASK_LMO_NPY and c0["icp_rel_fitness"] == thr and c0["icp_rel_rmse"] == thr and c0["zero_empty_centres"]
    assert c0["depth"].dtype == np.float64 and c0["ev"].symmetric is True
    assert res["n"] == 5 and res["evaluated"] == 3 and res["frames"] == stems and res["add_before"] == 3 / 5
    _FakeEvaluator.calls, _FakeEvaluator.empty_at = [], (0, 0)            # radii survive there: the reference would call Accumulator_3D on nothing
    with pytest.raises(ValueError):
        evaluate.evaluate_lmo_class(root, "glue", frames_per_batch=8, verbose=False)
    _FakeEvaluator.empty_at = None


# ---- checkpoint branch (AccumulatorSpace.py:594-610): the reference's own code run on stand-in network outputs ----
def _ckpt_dataset(tmp_path, g):
    n_frames, seed = (int(v) for v in g["lmckpt_cat_seed"])
    root = str(tmp_path) + "/"
    stems = synth.write_lm_dataset(root, "cat", n_frames, seed=seed)
    synth.write_lm_ckpt_maps(root, "cat", stems, seed=seed)
    return root, n_frames


def test_oracle_loop_matches_the_real_reference_checkpoint_branch(tmp_path, eval_golden):
    """CPU: mask rule `sem > 0.8 and radial <= max_radii` (a-1, :603-605) + cloud + vote + Horn + ADD as restated for the oracle
    against the reference's own checkpoint branch (its networks replaced by map files): survivors, keypoints (bit-identical),
    poses, ADD distance, ICP target size, ratio."""
    from rcvpose_b200 import evaluate
    root, n_frames = _ckpt_dataset(tmp_path, eval_golden)
    cls = evaluate.LinemodClass(root, "cat")
    assert cls.stems == list(eval_golden["lmckpt_cat_stems"])
    K = evaluate.linemod_K
    passed = 0
    for i, stem in enumerate(cls.stems):
        depth1 = cls.depth(stem)
        est, clouds = np.zeros((3, 3)), []
        for k in (1, 2, 3):
            sem, rad = synth.load_lm_ckpt_maps(root, "cat", k, stem)
            m = np.where(sem > 0.8, 1, 0)
            m = np.where(rad <= cls.max_radii_dm[k - 1], m, 0)
            dm = depth1 * m
            xyz_mm = oracle.rgbd_to_point_cloud(K, dm)
            rl = np.where(rad <= cls.max_radii_dm[k - 1], rad, 0)[dm.nonzero()]
            assert len(rl) == eval_golden["lmckpt_cat_n_points"][i, k - 1]
            est[k - 1] = oracle.Accumulator_3D(xyz_mm / 1000, rl)[0]
            clouds.append(xyz_mm)
        assert np.array_equal(est, eval_golden["lmckpt_cat_centres"][i])
        RT = np.zeros((4, 4))
        oracle.lmshorn(cls.keypoints_m[1:4] * 1000, est, 3, RT)
        np.testing.assert_allclose(RT, eval_golden["lmckpt_cat_RT"][i], rtol=0, atol=1e-9)
        gt = np.eye(4)
        gt[:3] = cls.pose_mm(stem)
        mean, _ = oracle.add_metric(cls.cad_m * 1000, RT, gt)
        assert abs(mean - eval_golden["lmckpt_cat_dist_before"][i]) <= 1e-9 * max(1.0, mean)
        assert len(oracle.scene_union(clouds)) == eval_golden["lmckpt_cat_scene_points"][i]
        passed += mean <= evaluate.add_threshold["cat"] * 1000
    assert passed / n_frames == eval_golden["lmckpt_cat_ratios"][0]


@_pending_gpu
def test_estimate_6d_pose_lm_checkpoint_branch_vs_real_reference_golden(tmp_path, eval_golden):
    """GPU: the drop-in evaluator's checkpoint branch (producer -> sem / radius maps -> MASK_LM_CKPT in rcv_vote_frames and
    rcv_scene_clouds) against what the reference's own checkpoint branch computed from the same maps."""
    from rcvpose_b200 import AccumulatorSpace as A
    root, n_frames = _ckpt_dataset(tmp_path, eval_golden)
    producer = lambda cls, k, path: synth.load_lm_ckpt_maps(root, cls, k, os.path.splitext(os.path.basename(path))[0])  # noqa: E731
    res = A.estimate_6d_pose_lm(types.SimpleNamespace(root_dataset=root, using_ckpts=True, producer=producer, classes=["cat"], frames_per_batch=2))["cat"]
    g = eval_golden
    assert res["frames"] == list(g["lmckpt_cat_stems"]) and np.array_equal(res["n_points"], g["lmckpt_cat_n_points"])
    assert np.array_equal(res["centre_mm"], g["lmckpt_cat_centres"])
    np.testing.assert_allclose(res["RT"], g["lmckpt_cat_RT"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(res["dist_before"], g["lmckpt_cat_dist_before"], rtol=1e-9, atol=0)
    assert np.array_equal(res["scene_points"], g["lmckpt_cat_scene_points"]) and res["add_before"] == g["lmckpt_cat_ratios"][0]


def test_ycb_format_readers(tmp_path):
    """N4, YCB-Video side: points.xyz and the -meta.mat fields the evaluator reads (AccumulatorSpace.py:988, :1015-1016, :1050-1057)."""
    import scipy.io
    rng = np.random.default_rng(4)
    pts = rng.normal(size=(25, 3)) * 0.05
    p = str(tmp_path / "points.xyz")
    np.savetxt(p, pts, fmt="%.9f")
    got = formats.read_xyz_points(p)
    assert got.shape == (25, 3) and np.allclose(got, pts, atol=1e-9)
    poses = rng.normal(size=(3, 4, 2))
    K = np.array([[1066.778, 0, 312.9869], [0, 1067.487, 241.3109], [0, 0, 1]])
    q = str(tmp_path / "000001-meta.mat")
    scipy.io.savemat(q, dict(intrinsic_matrix=K, factor_depth=np.array([[10000]], dtype=np.uint16), cls_indexes=np.array([[5], [12]], dtype=np.uint8),
                             poses=poses, center=np.zeros((2, 2))))
    meta = formats.load_ycb_meta(q)
    assert np.array_equal(meta["intrinsic_matrix"], K) and meta["factor_depth"] == 10000.0 and list(meta["cls_indexes"]) == [5, 12]
    assert meta["poses"].shape == (2, 3, 4) and np.array_equal(formats.pose_of(meta, 12), poses[:, :, 1]) and formats.pose_of(meta, 7) is None
    scipy.io.savemat(q, dict(intrinsic_matrix=K))
    with pytest.raises(ValueError):
        formats.load_ycb_meta(q)


@_pending_gpu
def test_lm_evaluator_revotes_items_beyond_max_grid(tmp_path):
    """a grid beyond the evaluator context's max_grid (D_EXCEEDS_CAP) no longer aborts the class -- the offending items
    are voted through the exact drop-in surface (grids up to 640) and the chain runs again with those centres.  With max_grid = 64
    every item of this dataset takes that route and must give the default run's results."""
    from rcvpose_b200 import evaluate
    root = str(tmp_path) + "/"
    synth.write_lm_dataset(root, "ape", 2, seed=9)
    a = evaluate.evaluate_lm_class(root, "ape", frames_per_batch=2, verbose=False)
    b = evaluate.evaluate_lm_class(root, "ape", frames_per_batch=2, verbose=False, max_grid=64)
    assert (a["grid"] > 64).all()
    for key in ("centre_mm", "scene_points"):
        assert np.array_equal(a[key], b[key]), key
    for key in ("RT", "dist_before", "RT_icp", "dist_after"):
        np.testing.assert_allclose(a[key], b[key], rtol=1e-12, atol=1e-9, err_msg=key)
    assert a["add_before"] == b["add_before"] and a["add_after"] == b["add_after"]
