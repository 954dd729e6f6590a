"""Synthetic LINEMOD-/YCB-shaped inputs for the voting path (no datasets are available offline).

The shapes follow SURVEY.md section 8d: a sphere-shaped object seen through the LINEMOD camera
(`linemod_K`, reference AccumulatorSpace.py:59-61), depth as uint16 millimetres (the `.dpt` format,
AccumulatorSpace.py:482-490), and one float32 radius map per keypoint in decimetres
(AccumulatorSpace.py:388: "radius map is in decimetre").
"""
import numpy as np

H, W = 480, 640

linemod_K = np.array([[572.4114, 0.0, 325.2611],
                      [0.0, 573.57043, 242.04899],
                      [0.0, 0.0, 1.0]])

ycb_K = np.array([[1066.778, 0.0, 312.9869],
                  [0.0, 1067.487, 241.3109],
                  [0.0, 0.0, 1.0]])


def sphere_depth(K, centre_mm, radius_mm, h=H, w=W):
    """Front ray-sphere intersection depth (mm, rounded to uint16); 0 where the ray misses."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    v, u = np.mgrid[0:h, 0:w].astype(np.float64)
    dx, dy = (u - cx) / fx, (v - cy) / fy          # ray = (dx, dy, 1) * z
    a = dx * dx + dy * dy + 1.0
    b = dx * centre_mm[0] + dy * centre_mm[1] + centre_mm[2]
    c = float(np.dot(centre_mm, centre_mm)) - radius_mm ** 2
    disc = b * b - a * c
    hit = disc > 0
    z = np.where(hit, (b - np.sqrt(np.where(hit, disc, 0.0))) / a, 0.0)
    return np.where(hit, np.round(z), 0).astype(np.uint16)


def backproject_mm(K, depth):
    """Host-side back-projection used only to *build* synthetic radius maps (same formula as
    AccumulatorSpace.py:77-85, dense form)."""
    h, w = depth.shape
    v, u = np.mgrid[0:h, 0:w].astype(np.float64)
    z = depth.astype(np.float64)
    x = ((u - K[0, 2]) * z) / K[0, 0]
    y = ((v - K[1, 2]) * z) / K[1, 1]
    return np.stack([x, y, z], axis=-1)


def radius_map_dm(K, depth, kpt_mm, rng, sigma_dm=0.01, outlier_frac=0.0, max_radius_dm=None):
    """float32 (H,W) radius map in decimetres: 10*|p-kpt| [m] + N(0,sigma); 0 outside the mask."""
    pts = backproject_mm(K, depth)
    dist_dm = np.linalg.norm(pts - kpt_mm[None, None, :], axis=-1) / 100.0
    r = dist_dm + rng.normal(0.0, sigma_dm, size=depth.shape)
    if outlier_frac > 0:
        out = rng.random(depth.shape) < outlier_frac
        hi = max_radius_dm if max_radius_dm is not None else float(dist_dm[depth != 0].max())
        r = np.where(out, rng.uniform(0.0, hi, size=depth.shape), r)
    return np.where(depth != 0, r, 0.0).astype(np.float32)


def config1_frame():
    """sphere r=50mm at (20,-10,900)mm, keypoint = centre+(80,70,60)mm,
    radius noise N(0,0.01 dm), default_rng(1234).  N=3189 px, D=86, 14,792,527 votes, peak 1215."""
    rng = np.random.default_rng(1234)
    centre = np.array([20.0, -10.0, 900.0])
    depth = sphere_depth(linemod_K, centre, 50.0)
    kpt = centre + np.array([80.0, 70.0, 60.0])
    radius = radius_map_dm(linemod_K, depth, kpt, rng)
    return dict(K=linemod_K, depth=depth, radius=radius[None], kpts_mm=kpt[None])


def config3_frame(f, n_kpts=3, K=linemod_K):
    """SURVEY 8d config 3 frame `f`: default_rng(f); object radius U(40,70)mm, centre x,y U(-150,150),
    z U(700,1100); keypoints at 1.5-2.5x object radius in dispersed directions; sigma 0.01 dm and
    2% uniform outliers in [0, max_radius]."""
    rng = np.random.default_rng(f)
    rad = rng.uniform(40.0, 70.0)
    centre = np.array([rng.uniform(-150, 150), rng.uniform(-150, 150), rng.uniform(700, 1100)])
    depth = sphere_depth(K, centre, rad)
    dirs = np.array([[1.0, 0.2, 0.1], [-0.3, 1.0, 0.2], [0.2, -0.4, 1.0], [-1.0, -0.5, 0.3]])[:n_kpts]
    dirs = dirs + rng.normal(0, 0.15, size=dirs.shape)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    kpts = centre[None] + dirs * rad * rng.uniform(1.5, 2.5, size=(n_kpts, 1))
    pts = backproject_mm(K, depth)
    radius, max_r = [], []
    for k in range(n_kpts):
        d = np.linalg.norm(pts - kpts[k][None, None], axis=-1) / 100.0
        mr = float(d[depth != 0].max()) if (depth != 0).any() else 1.0
        max_r.append(mr)
        radius.append(radius_map_dm(K, depth, kpts[k], rng, 0.01, 0.02, mr))
    return dict(K=K, depth=depth, radius=np.stack(radius), kpts_mm=kpts, max_radii_dm=np.array(max_r), obj_radius_mm=rad,
                centre_mm=centre)


def write_lm_dataset(root, class_name, n_frames, seed=0, obj_radius_mm=55.0, n_cad=1500, radius_dtype=np.float32, split_extra=2):
    """Lays out a synthetic class in the reference's LINEMOD directory structure (AccumulatorSpace.py:500-551, :566, :599, :612;
    see rcvpose_b200.evaluate.LinemodClass): a sphere-shaped object (CAD = points on the sphere, metres), `Outside9.npy`
    keypoints, and per frame a random pose, the rendered `.dpt` depth and three estimated radius maps (decimetres, noisy, 2 %
    outliers).  `split_extra` more images exist on disk than `Split/val.txt` lists.  Returns the list of test stems."""
    import os
    from . import formats
    rng = np.random.default_rng(seed)
    pv, orig = root + "LINEMOD/" + class_name + "/", root + "LINEMOD_ORIG/" + class_name + "/"
    for d in (pv + "JPEGImages", pv + "pose", pv + "Split", orig + "data"):
        os.makedirs(d, exist_ok=True)
    u = rng.normal(size=(n_cad, 3))
    cad_m = u / np.linalg.norm(u, axis=1, keepdims=True) * (obj_radius_mm / 1000)
    formats.write_ply_points(pv + class_name + ".ply", cad_m)
    cad_m = formats.read_ply_points(pv + class_name + ".ply")     # float32 on disk
    dirs = np.array([[0, 0, 0], [1.0, 0.2, 0.1], [-0.3, 1.0, 0.2], [0.2, -0.4, 1.0], [-1.0, -0.5, 0.3], [0.5, 0.5, -1.0], [1, 1, 1], [-1, 1, -1],
                     [1, -1, -1]], dtype=np.float64)
    dirs[1:] /= np.linalg.norm(dirs[1:], axis=1, keepdims=True)
    kp_m = dirs * (obj_radius_mm / 1000) * 1.8
    np.save(pv + "Outside9.npy", kp_m)
    max_r = [float(np.linalg.norm(cad_m - kp_m[k], axis=1).max() * 10) for k in (1, 2, 3)]
    stems = []
    for f in range(n_frames + split_extra):
        stem = "%06d" % (f * 7 + 3)
        open(pv + "JPEGImages/" + stem + ".jpg", "wb").close()
        if f >= n_frames:
            continue
        stems.append(stem)
        rv = rng.normal(size=3); th = np.linalg.norm(rv); k = rv / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t_m = np.array([rng.uniform(-0.15, 0.15), rng.uniform(-0.15, 0.15), rng.uniform(0.7, 1.1)])
        np.save(pv + "pose/pose" + str(int(stem)) + ".npy", np.concatenate([R, t_m[:, None]], axis=1))
        depth = sphere_depth(linemod_K, t_m * 1000, obj_radius_mm)
        formats.write_depth_dpt(orig + "data/depth" + str(int(stem)) + ".dpt", depth)
        for k in (1, 2, 3):
            d = root + "LINEMOD_ORIG/estRadialMap/" + class_name + "/Out_pt" + str(k) + "_dm/"
            os.makedirs(d, exist_ok=True)
            kpt_mm = (R @ kp_m[k] + t_m) * 1000
            r = radius_map_dm(linemod_K, depth, kpt_mm, rng, 0.01, 0.02, max_r[k - 1] * 1.2)   # some outliers exceed max_radii
            np.save(d + stem + ".npy", r.astype(radius_dtype))
    with open(pv + "Split/val.txt", "w") as fh:
        fh.write("".join(s + "\n" for s in stems))
    return stems


def write_lm_ckpt_maps(root, class_name, stems, seed=0):
    """Stand-ins for the outputs of the three keypoint networks on a class written by write_lm_dataset (the reference's checkpoint
    branch, AccumulatorSpace.py:594-610): per (keypoint k, frame) a seg score map (0.95 on most of the object, 0.6 on the rest of
    it, 0.1 elsewhere -- never above the 0.8 threshold where depth is 0, which would trip the reference's list mis-alignment,
    SURVEY 8a a-2) and a radius map that, like a network's, has values everywhere (the estimated radii on the object, noise off it).
    Files: <root>ckpt_maps/<cls>/pt<k>/<stem>_sem.npy, _radial.npy (float32)."""
    import os
    from . import formats
    rng = np.random.default_rng(seed)
    for stem in stems:
        depth = formats.read_depth(root + "LINEMOD_ORIG/" + class_name + "/data/depth" + str(int(stem)) + ".dpt")
        for k in (1, 2, 3):
            d = root + "ckpt_maps/" + class_name + "/pt" + str(k) + "/"
            os.makedirs(d, exist_ok=True)
            est = np.load(os.path.join(root + "LINEMOD_ORIG/", "estRadialMap", class_name, "Out_pt" + str(k) + "_dm", stem + ".npy")).astype(np.float32)
            obj = depth != 0
            sem = np.where(obj, np.where(rng.random(depth.shape) < 0.85, 0.95, 0.6), 0.1).astype(np.float32)
            radial = np.where(obj, est, rng.uniform(0.0, 3.0, size=depth.shape)).astype(np.float32)
            np.save(d + stem + "_sem.npy", sem)
            np.save(d + stem + "_radial.npy", radial)


def load_lm_ckpt_maps(root, class_name, k, stem):
    d = root + "ckpt_maps/" + class_name + "/pt" + str(k) + "/"
    return np.load(d + stem + "_sem.npy"), np.load(d + stem + "_radial.npy")


def write_lmo_dataset(root, class_name, n_frames, seed=0, obj_radius_mm=55.0, n_cad=1500, radius_dtype=np.float32):
    """Synthetic class in the reference's Occlusion-LINEMOD layout (AccumulatorSpace.py:746-850; rcvpose_b200.evaluate.LmoClass).
    Besides n_frames complete frames the image directory holds one frame without a pose and one without the third radius map
    (both are counted but not evaluated, :962; a non-image entry would crash the reference at :896); in the second complete frame the radius map of
    keypoint 2 is all zero (the reference skips that keypoint, :858).  Returns the stems of the complete frames."""
    import os
    from PIL import Image
    from . import formats
    rng = np.random.default_rng(seed)
    pv, occ = root + "LINEMOD/" + class_name + "/", root + "OCCLUSION_LINEMOD/"
    for d in (pv, occ + "RGB-D/rgb_noseg", occ + "RGB-D/depth_noseg", occ + "blender_poses/" + class_name):
        os.makedirs(d, exist_ok=True)
    u = rng.normal(size=(n_cad, 3))
    formats.write_ply_points(pv + class_name + ".ply", u / np.linalg.norm(u, axis=1, keepdims=True) * (obj_radius_mm / 1000))
    cad_m = formats.read_ply_points(pv + class_name + ".ply")
    dirs = np.array([[0, 0, 0], [1.0, 0.2, 0.1], [-0.3, 1.0, 0.2], [0.2, -0.4, 1.0], [-1.0, -0.5, 0.3], [0.5, 0.5, -1.0], [1, 1, 1], [-1, 1, -1],
                     [1, -1, -1]], dtype=np.float64)
    dirs[1:] /= np.linalg.norm(dirs[1:], axis=1, keepdims=True)
    kp_m = dirs * (obj_radius_mm / 1000) * 1.8
    np.save(pv + "Outside9.npy", kp_m)
    max_r = [float(np.linalg.norm(cad_m - kp_m[k], axis=1).max() * 10) for k in (1, 2, 3)]
    stems = []
    for f in range(n_frames + 2):
        idx = f * 5 + 2
        stem = "color_%05d" % idx
        open(occ + "RGB-D/rgb_noseg/" + stem + ".png", "wb").close()
        rv = rng.normal(size=3); th = np.linalg.norm(rv); k = rv / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t_m = np.array([rng.uniform(-0.15, 0.15), rng.uniform(-0.15, 0.15), rng.uniform(0.7, 1.1)])
        if f != n_frames:                                  # frame n_frames has no pose
            np.save(occ + "blender_poses/" + class_name + "/pose" + str(idx) + ".npy", np.concatenate([R, t_m[:, None]], axis=1))
        depth = sphere_depth(linemod_K, t_m * 1000, obj_radius_mm)
        Image.fromarray(depth).save(occ + "RGB-D/depth_noseg/depth_%05d.png" % idx)
        for k in (1, 2, 3):
            if f == n_frames + 1 and k == 3:               # frame n_frames + 1 lacks the third map
                continue
            d = occ + "estRadialMap/" + class_name + "/Out_pt" + str(k) + "_dm/"
            os.makedirs(d, exist_ok=True)
            r = radius_map_dm(linemod_K, depth, (R @ kp_m[k] + t_m) * 1000, rng, 0.01, 0.02, max_r[k - 1] * 1.2)
            if f == 1 and k == 2:
                r = r * 0 + 9.0e3 * (depth != 0)           # everything beyond max_radii: the thresholded map is all zero
            np.save(d + "_%05d.npy" % idx, r.astype(radius_dtype))
        if f < n_frames:
            stems.append(stem)
    return stems


def frame_to_points(K, depth, radius):
    """The reference caller's glue (AccumulatorSpace.py:612-619, npy branch): masked depth ->
    xyz in metres (N,3) float64 + radial list (N,) in the map's dtype."""
    m = (radius != 0)
    dm = depth * m
    vs, us = dm.nonzero()
    zs = dm[vs, us]
    xs = ((us - K[0, 2]) * zs) / float(K[0, 0])
    ys = ((vs - K[1, 2]) * zs) / float(K[1, 1])
    xyz_mm = np.array([xs, ys, zs]).T
    return xyz_mm / 1000, radius[dm.nonzero()]


def frame_params(n_frames, n_kpts=3, seed=0, obj_radius_mm=(40.0, 70.0), z_mm=(700.0, 1100.0), approach=False):
    """Scenes of a GLOBAL frame sequence (host, NumPy), to be sliced per rank and rendered with torch_batch(params=...): object
    radius, centre and the keypoints in the object frame, drawn like torch_batch draws them.  approach=True orders the frames
    by decreasing distance (a camera approaching the object), so that the cost of a frame grows along the sequence."""
    rng = np.random.default_rng(seed)
    obj_r = rng.uniform(obj_radius_mm[0], obj_radius_mm[1], n_frames)
    z = rng.uniform(z_mm[0], z_mm[1], n_frames)
    if approach:
        z = np.sort(z)[::-1].copy()
    centre = np.stack([rng.uniform(-150, 150, n_frames), rng.uniform(-150, 150, n_frames), z], axis=1)
    base = np.array([[1.0, 0.2, 0.1], [-0.3, 1.0, 0.2], [0.2, -0.4, 1.0], [-1.0, -0.5, 0.3]])[:n_kpts]
    dirs = base[None] + 0.15 * rng.normal(size=(n_frames, n_kpts, 3))
    dirs /= np.linalg.norm(dirs, axis=2, keepdims=True)
    model = dirs * (obj_r[:, None, None] * (1.5 + rng.random((n_frames, n_kpts, 1))))
    return dict(obj_r=obj_r, centre=centre, model=model)


def frame_cost(params, K=linemod_K, acc_unit_mm=5.0):
    """The cost model of SURVEY.md 8e for a frame: sum over keypoints of N * R^2 -- N = surviving pixels (the projected disc of the
    object), R = the keypoint's mean distance to the visible surface in voxels.  Votes of a point grow like 4 pi R^2 f, the
    rasteriser's work like pi R^2 + 2 R columns."""
    r, z = params["obj_r"], params["centre"][:, 2]
    n_px = np.pi * (K[0, 0] * r / z) * (K[1, 1] * r / z)
    R = np.linalg.norm(params["model"], axis=2) / acc_unit_mm
    return (n_px[:, None] * R * R).sum(axis=1)


def torch_batch(n_frames, n_kpts=3, seed=0, device="cuda", K=linemod_K, chunk=128, sigma_dm=0.01, outlier_frac=0.02, h=H, w=W,
                obj_radius_mm=(40.0, 70.0), params=None):
    """Config-3 shaped batch generated on the GPU: per frame an object sphere of radius
    U(obj_radius_mm) = U(40,70) mm at x,y U(-150,150), z U(700,1100) mm; `n_kpts` keypoints at 1.5-2.5 object radii in
    dispersed directions; radius maps (decimetres, float32) with N(0, sigma) noise and a fraction of
    uniform outliers in [0, max radius]; depth uint16 millimetres (returned as an int16 view).
    Returns dict(depth (B,H,W) int16-view-of-uint16, radius (B,Kp,H,W) f32, kpts_mm (B,Kp,3) f64,
    centre_mm (B,3) f64, model_mm (B,Kp,3) f64 = keypoints in the object frame)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    dev = torch.device(device)
    depth = torch.empty((n_frames, h, w), dtype=torch.int16, device=dev)
    radius = torch.empty((n_frames, n_kpts, h, w), dtype=torch.float32, device=dev)
    U = lambda *s: torch.rand(*s, generator=g, device=dev, dtype=torch.float64)  # noqa: E731
    if params is not None:     # the scene of every frame is given (frame_params: a rank's slice of a global sequence)
        obj_r = torch.as_tensor(params["obj_r"], dtype=torch.float64, device=dev)
        centre = torch.as_tensor(params["centre"], dtype=torch.float64, device=dev)
        model = torch.as_tensor(params["model"], dtype=torch.float64, device=dev)
        assert obj_r.shape[0] == n_frames and model.shape[1] == n_kpts
    else:
        obj_r = obj_radius_mm[0] + (obj_radius_mm[1] - obj_radius_mm[0]) * U(n_frames)
        centre = torch.stack([-150 + 300 * U(n_frames), -150 + 300 * U(n_frames), 700 + 400 * U(n_frames)], dim=1)
        base = torch.tensor([[1.0, 0.2, 0.1], [-0.3, 1.0, 0.2], [0.2, -0.4, 1.0], [-1.0, -0.5, 0.3]], dtype=torch.float64, device=dev)[:n_kpts]
        dirs = base[None] + 0.15 * torch.randn((n_frames, n_kpts, 3), generator=g, device=dev, dtype=torch.float64)
        dirs = dirs / dirs.norm(dim=2, keepdim=True)
        model = dirs * (obj_r[:, None, None] * (1.5 + U(n_frames, n_kpts, 1)))
    kpts = centre[:, None, :] + model
    fx, fy, cx, cy = float(K[0, 0]), float(K[1, 1]), float(K[0, 2]), float(K[1, 2])
    vv, uu = torch.meshgrid(torch.arange(h, device=dev, dtype=torch.float64), torch.arange(w, device=dev, dtype=torch.float64), indexing="ij")
    dx, dy = (uu - cx) / fx, (vv - cy) / fy
    a = dx * dx + dy * dy + 1.0
    for f0 in range(0, n_frames, chunk):
        f1 = min(n_frames, f0 + chunk)
        c = centre[f0:f1]
        b = dx[None] * c[:, 0, None, None] + dy[None] * c[:, 1, None, None] + c[:, 2, None, None]
        cc = (c * c).sum(dim=1) - obj_r[f0:f1] ** 2
        disc = b * b - a[None] * cc[:, None, None]
        hit = disc > 0
        z = torch.where(hit, (b - torch.sqrt(disc.clamp_min(0))) / a[None], torch.zeros_like(b)).round()
        depth[f0:f1] = z.to(torch.int32).to(torch.int16)          # values < 32768: the int16 view equals the uint16 value
        x = (uu[None] - cx) * z / fx
        y = (vv[None] - cy) * z / fy
        for k in range(n_kpts):
            kp = kpts[f0:f1, k]
            d = torch.sqrt((x - kp[:, 0, None, None]) ** 2 + (y - kp[:, 1, None, None]) ** 2 + (z - kp[:, 2, None, None]) ** 2) / 100.0
            mr = torch.where(hit, d, torch.zeros_like(d)).amax(dim=(1, 2))
            r = d + sigma_dm * torch.randn(d.shape, generator=g, device=dev, dtype=torch.float64)
            out = torch.rand(d.shape, generator=g, device=dev) < outlier_frac
            r = torch.where(out, mr[:, None, None] * torch.rand(d.shape, generator=g, device=dev, dtype=torch.float64), r)
            radius[f0:f1, k] = torch.where(hit, r, torch.zeros_like(r)).to(torch.float32)
    return dict(depth=depth, radius=radius, kpts_mm=kpts, centre_mm=centre, model_mm=model.contiguous())


# ------------------------------------------------------------------------------------------------
# YCB-Video-shaped class in the layout the reference's estimate_6d_pose_ycb reads (AccumulatorSpace.py:981-1057)
# ------------------------------------------------------------------------------------------------
def write_ycb_dataset(root, class_id, class_name, n_frames, seed=0, obj_radius_mm=55.0, n_cad=1200, cycle="0048", split_extra=1):
    """<root>/Split/<cls>/val.txt ("<cycle>_<idx>" per line), models/<cls>/{points.xyz, Outside9.npy}, and per frame
    data/<cycle>/<idx>.mat (the path the reference hands scipy.io.loadmat, :1015: poses (3,4,M) in metres, cls_indexes (M,1),
    factor_depth, intrinsic_matrix), <idx>-depth.png (uint16, depth * factor_depth) and <idx>-color.png.  Every frame holds the
    class's object (a sphere, like write_lm_dataset) and one more object of another class, so that the depth image is non-zero
    off the object and cls_indexes has to be searched.  Returns the frame names listed in val.txt."""
    import os
    import scipy.io
    from PIL import Image
    rng = np.random.default_rng(seed)
    md = root + "models/" + class_name + "/"
    for d in (md, root + "Split/" + class_name, root + "data/" + cycle):
        os.makedirs(d, exist_ok=True)
    u = rng.normal(size=(n_cad, 3))
    cad_m = u / np.linalg.norm(u, axis=1, keepdims=True) * (obj_radius_mm / 1000) * np.array([1.0, 0.8, 0.6])   # an ellipsoid cloud: the OBB is not a cube
    np.savetxt(md + "points.xyz", cad_m, fmt="%.8f")
    dirs = np.array([[0, 0, 0], [1.0, 0.2, 0.1], [-0.3, 1.0, 0.2], [0.2, -0.4, 1.0], [-1.0, -0.5, 0.3], [0.5, 0.5, -1.0], [1, 1, 1], [-1, 1, -1],
                     [1, -1, -1]], dtype=np.float64)
    dirs[1:] /= np.linalg.norm(dirs[1:], axis=1, keepdims=True)
    kp_m = dirs * (obj_radius_mm / 1000) * 1.8
    np.save(md + "Outside9.npy", kp_m)
    names = []
    other_id = 1 if class_id != 1 else 2
    for f in range(n_frames + split_extra):
        idx = "%06d" % (f * 5 + 1)
        name = cycle + "_" + idx
        rv = rng.normal(size=3); th = np.linalg.norm(rv); k = rv / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        t_m = np.array([rng.uniform(-0.08, 0.02), rng.uniform(-0.06, 0.06), rng.uniform(0.85, 1.1)])
        t2_m = t_m + np.array([0.2, 0.02, 0.05])                                     # the other object, to the right
        Kf = ycb_K.copy()
        Kf[0, 2] += f                                                                # intrinsics are per frame (:1057)
        d1 = sphere_depth(Kf, t_m * 1000, obj_radius_mm).astype(np.float64)
        d2 = sphere_depth(Kf, t2_m * 1000, 40.0).astype(np.float64)
        depth_mm = np.where(d1 > 0, d1, d2)
        factor = 10000.0
        Image.fromarray(np.round(depth_mm * (factor / 1000)).astype(np.uint16)).save(root + "data/" + cycle + "/" + idx + "-depth.png")
        Image.fromarray(np.zeros((H, W, 3), np.uint8)).save(root + "data/" + cycle + "/" + idx + "-color.png")
        poses = np.stack([np.concatenate([np.eye(3), t2_m[:, None]], axis=1), np.concatenate([R, t_m[:, None]], axis=1)], axis=2)
        scipy.io.savemat(root + "data/" + cycle + "/" + idx + ".mat",
                         dict(poses=poses, cls_indexes=np.array([[other_id], [class_id]], dtype=np.uint8), factor_depth=np.array([[factor]]),
                              intrinsic_matrix=Kf))
        if f < n_frames:
            names.append(name)
    with open(root + "Split/" + class_name + "/val.txt", "w") as fh:
        fh.write("".join(s + "\n" for s in names))
    return names


def write_ycb_ckpt_maps(root, class_id, class_name, names, seed=0):
    """Stand-ins for the three keypoint networks' outputs on a class written by write_ycb_dataset: per (keypoint k = 1..3, frame) a
    seg score map (0.95 on ~40 % of the object's pixels, 0.6 on the rest of it, 0.1 elsewhere -- including the OTHER object, where
    depth is non-zero) and a radius map in decimetres with values everywhere.  <root>ckpt_maps/<cls>/pt<k>/<name>_sem.npy, _radial.npy."""
    import os
    import scipy.io
    from PIL import Image
    rng = np.random.default_rng(seed + 1)
    kp_m = np.load(root + "models/" + class_name + "/Outside9.npy")
    for name in names:
        cycle, idx = name.split("_")
        meta = scipy.io.loadmat(root + "data/" + cycle + "/" + idx + ".mat")
        Kf = meta["intrinsic_matrix"]
        j = int(np.nonzero(np.ravel(meta["cls_indexes"]) == class_id)[0][0])
        RT = meta["poses"][:, :, j]
        depth_m = np.array(Image.open(root + "data/" + cycle + "/" + idx + "-depth.png")).astype(np.float64) / float(meta["factor_depth"][0, 0])
        obj = sphere_depth(Kf, RT[:, 3] * 1000, 1.0e-3 + float(np.linalg.norm(np.load(root + "models/" + class_name + "/Outside9.npy")[1]) / 1.8 * 1000)) > 0
        obj &= depth_m > 0
        for k in (1, 2, 3):
            d = root + "ckpt_maps/" + class_name + "/pt" + str(k) + "/"
            os.makedirs(d, exist_ok=True)
            kpt_mm = (RT[:, :3] @ kp_m[k] + RT[:, 3]) * 1000
            est = radius_map_dm(Kf, np.where(obj, depth_m * 1000, 0.0), kpt_mm, rng, 0.01, 0.02, None)
            sem = np.where(obj, np.where(rng.random(obj.shape) < 0.4, 0.95, 0.6), 0.1).astype(np.float32)
            radial = np.where(obj, est, rng.uniform(0.0, 3.0, size=obj.shape)).astype(np.float32)
            np.save(d + name + "_sem.npy", sem)
            np.save(d + name + "_radial.npy", radial)


def load_ycb_ckpt_maps(root, class_name, k, name):
    d = root + "ckpt_maps/" + class_name + "/pt" + str(k) + "/"
    return np.load(d + name + "_sem.npy"), np.load(d + name + "_radial.npy")
