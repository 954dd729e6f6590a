"""GPU parity tests: the CUDA path (through the C ABI, librcvvote.so) against the CPU oracle and the
golden fixtures produced by the real reference.  Integer results (vote volumes, peaks, D, zero
boundary, point counts) must be bit-exact; float64 centres are compared exactly where the operation
order is reproduced (they are affine in exact integers and the numpy-pairwise mean) and Horn poses to
1e-12 absolute on rotation entries (north-star tolerance: 1e-4 relative)."""
import numpy as np
import pytest
import torch

from oracle import oracle
from rcvpose_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def acc():
    from rcvpose_b200 import AccumulatorSpace as A
    return A


@pytest.fixture(scope="module")
def ctx():
    from rcvpose_b200 import api
    return api.VoteContext(0, max_items=64, max_points_total=1 << 21, max_grid=400)


def test_library_loaded_and_device_is_sm100():
    from rcvpose_b200 import _lib
    L = _lib.load()
    assert L.rcv_abi_version() == _lib.RCV_ABI_VERSION
    assert torch.cuda.get_device_capability(0)[0] == 10


def test_rgbd_to_point_cloud_exact(acc, golden):
    fr = synth.config1_frame()
    dm = fr["depth"] * (fr["radius"][0] != 0)
    got = acc.rgbd_to_point_cloud(fr["K"], dm)
    assert got.dtype == np.float64 and np.array_equal(got, golden["c1_xyz_mm"])
    got64 = acc.rgbd_to_point_cloud(fr["K"], dm.astype(np.float64))
    assert np.array_equal(got64, golden["c1_xyz_mm"])


def test_config1_volume_bit_exact_vs_reference(acc, golden):
    fr = synth.config1_frame()
    xyz, rl = synth.frame_to_points(fr["K"], fr["depth"], fr["radius"][0])
    vol, out = acc.vote_volume(xyz, rl)
    assert vol.shape == (86, 86, 86)
    assert np.array_equal(vol, golden["c1_volume"])            # every voxel equals the reference's VoteMap_3D
    assert int(out["votes"].item()) == int(golden["c1_votes"])
    assert int(out["peak"].item()) == int(golden["c1_peak"])
    assert int(out["zero_boundary"].item()) == int(golden["c1_zb"])
    centre, info = acc.Accumulator_3D(xyz, rl, return_info=True)
    assert centre.shape == (1, 3) and centre.dtype == np.float64
    assert np.array_equal(centre[0], golden["c1_centre_mm"])   # same float64 bits as the reference
    assert info["D"] == 86 and info["votes"] == int(golden["c1_votes"])


@pytest.mark.parametrize("t", range(8))
def test_random_clouds_vs_reference(acc, golden, t):
    xyz, rl = golden["r%d_xyz" % t], golden["r%d_rl" % t]
    vol, out = acc.vote_volume(xyz, rl)
    assert np.array_equal(vol, golden["r%d_volume" % t])
    assert int(out["grid"].item()) == int(golden["r%d_D" % t]) and int(out["zero_boundary"].item()) == int(golden["r%d_zb" % t])
    c = acc.Accumulator_3D(xyz, rl)
    assert np.array_equal(c[0], golden["r%d_centre" % t][0])


def test_frames_api_vs_reference(ctx, golden):
    frames = [synth.config3_frame(f) for f in range(4)]
    depth = torch.from_numpy(np.stack([f["depth"] for f in frames]).view(np.int16)).cuda()
    radius = torch.from_numpy(np.stack([f["radius"] for f in frames])).cuda()
    K = torch.from_numpy(synth.linemod_K).cuda()
    from rcvpose_b200 import api
    out = ctx.vote_frames(depth, radius, K, mask_flags=api.RCV_MASK_RADIUS_NONZERO)
    torch.cuda.synchronize()
    for f in range(4):
        for k in range(3):
            g = lambda name: golden["c3_f%d_k%d_%s" % (f, k, name)]
            assert int(out["status"][f, k]) == 0
            assert int(out["n_points"][f, k]) == int(g("n"))
            assert int(out["grid"][f, k]) == int(g("D"))
            assert int(out["votes"][f, k]) == int(g("votes"))
            assert int(out["peak"][f, k]) == int(g("peak"))
            assert np.array_equal(out["centre_mm"][f, k].cpu().numpy(), g("centre_mm"))
    # host-buffer entry point gives the same answers
    h = ctx.vote_frames_host(np.stack([f["depth"] for f in frames]), np.stack([f["radius"] for f in frames]), synth.linemod_K,
                             mask_flags=api.RCV_MASK_RADIUS_NONZERO, frames_per_chunk=3)
    for key in ("centre_mm", "peak", "votes", "n_points", "grid", "status"):
        assert np.array_equal(h[key], out[key].cpu().numpy()), key


def test_host_entry_point_large_batch_equals_device_path():
    """rcv_vote_frames_host at a batch size that takes every branch of the host side -- helper threads scanning the depth rows
    ahead of the copies (n_frames >= 64), the ramped first chunks (n_frames >= 4 chunks of >= 128), the strided copy of a
    frame's keypoint planes, frames with empty / full-height / single-row depth -- against rcv_vote_frames on device-resident
    copies of the same arrays: every output identical (the row cropping is bit-neutral), and fewer bytes crossed the bus."""
    from rcvpose_b200 import api
    B, Kp, H, W = 640, 2, 48, 64
    rng = np.random.default_rng(17)
    depth = np.zeros((B, H, W), np.uint16)
    radius = rng.uniform(0.25, 0.6, size=(B, Kp, H, W)).astype(np.float32)
    sem = rng.uniform(0.0, 1.0, size=(B, Kp, H, W)).astype(np.float32)
    for f in range(B):
        kind = f % 7
        if kind == 0:
            continue                                                   # no depth at all: RCV_ST_EMPTY_MASK
        r0 = int(rng.integers(0, H - 1)); r1 = int(rng.integers(r0 + 1, min(H, r0 + 20) + 1))
        if kind == 1:
            r0, r1 = 0, H                                              # full height: nothing to crop
        if kind == 2:
            r1 = r0 + 1                                                # a single row
        c0 = int(rng.integers(0, W - 8)); c1 = int(rng.integers(c0 + 4, min(W, c0 + 24) + 1))
        depth[f, r0:r1, c0:c1] = rng.integers(700, 1100, size=(r1 - r0, c1 - c0)).astype(np.uint16)
        if kind == 3:
            depth[f, r0, c0:c1] = 0                                    # first row of the block empty again
    K = np.array([[57.3, 0.0, 32.5], [0.0, 57.4, 24.2], [0.0, 0.0, 1.0]])
    mr = np.full((Kp,), 0.55, np.float64)
    flags = api.RCV_MASK_MAX_RADIUS | api.RCV_MASK_SEM_GT
    ctx = api.VoteContext(0, max_items=256 * Kp, max_points_total=1 << 22, max_grid=256, image=(H, W))
    h = ctx.vote_frames_host(depth, radius, K, sem=sem, max_radii=mr, mask_flags=flags, sem_threshold=0.3, frames_per_chunk=128)
    sent = ctx.last_h2d_bytes
    outs = []
    for f0 in range(0, B, 256):                                        # the device path, in batches the context holds
        sl = slice(f0, min(B, f0 + 256))
        o = ctx.vote_frames(torch.from_numpy(depth[sl].view(np.int16)).cuda(), torch.from_numpy(radius[sl]).cuda(), torch.from_numpy(K).cuda(),
                            sem=torch.from_numpy(sem[sl]).cuda(), max_radii=torch.from_numpy(mr).cuda(), mask_flags=flags, sem_threshold=0.3)
        torch.cuda.synchronize()
        outs.append({k: v.cpu().numpy() for k, v in o.items() if k in ("centre_mm", "peak", "votes", "n_points", "grid", "status")})
    for key in ("centre_mm", "peak", "votes", "n_points", "grid", "status"):
        want = np.concatenate([o[key] for o in outs])
        assert np.array_equal(h[key], want), key
    assert (h["status"][::7] == api.RCV_ST_EMPTY_MASK).all() and (h["status"] == 0).sum() > B * Kp // 2
    assert 0 < sent < depth.nbytes + radius.nbytes + sem.nbytes, sent


def test_mask_rules_and_max_radius(ctx):
    from rcvpose_b200 import api
    fr = synth.config3_frame(11)
    rng = np.random.default_rng(3)
    sem = (rng.random(fr["radius"].shape) * 1.2).astype(np.float32)
    maxr = fr["max_radii_dm"] * 0.9
    depth = torch.from_numpy(fr["depth"][None].view(np.int16)).cuda()
    radius = torch.from_numpy(fr["radius"][None]).cuda()
    out = ctx.vote_frames(depth, radius, torch.from_numpy(fr["K"]).cuda(), sem=torch.from_numpy(sem[None]).cuda(),
                          max_radii=torch.from_numpy(maxr).cuda(), mask_flags=api.MASK_LM_CKPT, sem_threshold=0.8)
    torch.cuda.synchronize()
    for k in range(3):
        # reference rule AccumulatorSpace.py:603-610 with "surviving pixel = mask && depth != 0" 
        m = (sem[k] > 0.8) & (fr["radius"][k] <= maxr[k]) & (fr["depth"] != 0)
        xyz_mm = oracle.rgbd_to_point_cloud(fr["K"], fr["depth"] * m)
        rl = fr["radius"][k][m]
        want, info = oracle.Accumulator_3D(xyz_mm / 1000, rl, method="scatter", return_info=True)
        assert int(out["n_points"][0, k]) == int(m.sum())
        assert int(out["grid"][0, k]) == info["D"] and int(out["votes"][0, k]) == info["votes"] and int(out["peak"][0, k]) == info["peak"]
        assert np.array_equal(out["centre_mm"][0, k].cpu().numpy(), want[0])


def test_float64_radii_and_ycbgen_policy(acc):
    rng = np.random.default_rng(5)
    xyz = rng.normal(0, 0.03, size=(300, 3)) + np.array([0.1, -0.2, 0.8])
    kp = xyz.mean(0) + np.array([0.05, 0.02, -0.04])
    rl = np.linalg.norm(xyz - kp, axis=1) * 10
    for policy, scale, r in ((0, 100, rl), (1, 1000, rl / 10)):
        want, info = oracle.Accumulator_3D(xyz, r, radius_scale=scale, policy=policy, return_info=True, return_volume=True)
        vol, out = acc.vote_volume(xyz, r, radius_scale=scale, policy=policy)
        assert np.array_equal(vol, info["volume"])
        got = acc.Accumulator_3D(xyz, r, radius_scale=scale, policy=policy)
        assert np.array_equal(got, want)


def test_large_grid_uses_row_tiles(acc):
    # acc_unit 1.25 mm -> D ~ 300: one slice no longer fits a CTA's shared memory (ni = 1, nj < D)
    rng = np.random.default_rng(8)
    xyz = rng.normal(0, 0.025, size=(400, 3)) + np.array([0.0, 0.1, 0.9])
    kp = xyz.mean(0) + np.array([0.12, 0.05, -0.08])
    rl = (np.linalg.norm(xyz - kp, axis=1) * 10).astype(np.float32)
    want, info = oracle.Accumulator_3D(xyz, rl, acc_unit=1.25, method="scatter", return_info=True, return_volume=True)
    assert info["D"] > 240
    vol, out = acc.vote_volume(xyz, rl, acc_unit=1.25)
    assert np.array_equal(vol, info["volume"])
    assert np.array_equal(acc.Accumulator_3D(xyz, rl, acc_unit=1.25), want)


def test_ycb_shaped_frame_both_grid_policies():
    """BASELINE configs[3]: YCB-Video geometry (camera, uint16 depth with factor_depth 10000, metres), a large object so that
    N ~ 24 k points, R up to ~100 voxels (outliers), D ~ 235: the frames entry point against the oracle fed the same maps, LM grid policy
    (what estimate_6d_pose_ycb calls, AccumulatorSpace.py:1067) and the YCBGEN policy (3DRadius_ycb.py:113-141, metre radii)."""
    from rcvpose_b200 import api
    K = synth.ycb_K
    centre = np.array([40.0, -25.0, 820.0])
    depth_mm = synth.sphere_depth(K, centre, 95.0)                        # uint16 millimetres
    kpt_mm = centre + np.array([170.0, 110.0, -130.0])
    rng = np.random.default_rng(77)
    radius_dm = synth.radius_map_dm(K, depth_mm, kpt_mm, rng, sigma_dm=0.01, outlier_frac=0.01, max_radius_dm=5.0)
    vv, uu = np.indices(radius_dm.shape)
    radius_dm[(vv + uu) % 2 == 1] = 0.0                                   # checkerboard mask: keeps the oracle's run time in seconds
    depth_raw = (depth_mm.astype(np.uint32) * 10).astype(np.uint16)          # factor_depth = 10000 -> metres
    ctx = api.VoteContext(0, max_items=2, max_points_total=1 << 17, max_grid=320)
    Kd = torch.from_numpy(K).cuda()
    d_t = torch.from_numpy(depth_raw.view(np.int16))[None].cuda()
    for policy, scale, rmap in ((api.RCV_POLICY_LM, 100.0, radius_dm), (api.RCV_POLICY_YCBGEN, 1000.0, (radius_dm / 10).astype(np.float32))):
        out = ctx.vote_frames(d_t, torch.from_numpy(rmap)[None, None].cuda(), Kd, mask_flags=api.RCV_MASK_RADIUS_NONZERO, depth_div=10000.0,
                              xyz_div=1.0, radius_scale=scale, policy=policy)
        torch.cuda.synchronize()
        m = (rmap != 0) & (depth_raw != 0)
        xyz = oracle.rgbd_to_point_cloud(K, np.where(m, depth_raw / 10000.0, 0.0))
        rl = rmap[np.nonzero(np.where(m, depth_raw, 0))]
        want, info = oracle.Accumulator_3D(xyz, rl, radius_scale=scale, policy=policy, method="scatter", return_info=True)
        assert int(out["status"][0, 0]) == 0
        assert int(out["n_points"][0, 0]) == xyz.shape[0] and xyz.shape[0] > 20000
        assert int(out["grid"][0, 0]) == info["D"] and info["D"] > 120
        assert int(out["votes"][0, 0]) == info["votes"]
        assert int(out["peak"][0, 0]) == info["peak"]
        assert np.array_equal(out["centre_mm"][0, 0].cpu().numpy(), want[0])


def test_fine_voxel_stress_grid_512(acc):
    """BASELINE configs[4]: accumulator resolution sweep up to a ~512^3 grid (acc_unit 1.2 mm): a slice no longer fits one CTA's
    shared memory, tiles are single slices x row bands and the volume lives in HBM only for this parity dump."""
    rng = np.random.default_rng(11)
    n = 400
    xyz = rng.normal(0, 0.02, size=(n, 3)) + np.array([0.05, -0.02, 0.85])
    kp = xyz.mean(0) + np.array([0.11, 0.06, -0.09])
    rl = (np.linalg.norm(xyz - kp, axis=1) * 10 + rng.normal(0, 0.005, n)).astype(np.float32)
    for unit in (2.5, 1.2):
        want, info = oracle.Accumulator_3D(xyz, rl, acc_unit=unit, method="scatter", return_info=True, return_volume=True)
        got, ginfo = acc.Accumulator_3D(xyz, rl, acc_unit=unit, return_info=True)
        assert ginfo["D"] == info["D"] and ginfo["votes"] == info["votes"] and ginfo["peak"] == info["peak"]
        assert np.array_equal(got, want)
        if unit == 1.2:
            assert 440 < info["D"] <= 640
            vol, out = acc.vote_volume(xyz, rl, acc_unit=unit)
            assert np.array_equal(vol, info["volume"])


def test_empty_and_degenerate_inputs(acc, ctx):
    with pytest.raises(ValueError):
        acc.Accumulator_3D(np.zeros((0, 3)), np.zeros((0,), np.float32))
    # all radii negative: R <= 0 never votes; the reference returns voxel (0,0,0) un-shifted (all-equal volume)
    xyz = np.array([[0.0, 0.0, 0.5], [0.01, 0.0, 0.5], [0.0, 0.02, 0.52], [0.3, 0.1, 0.4]])
    rl = np.array([0.01, 0.02, 0.0, 0.03], np.float32)
    want, info = oracle.Accumulator_3D(xyz, rl, return_info=True)
    got, ginfo = acc.Accumulator_3D(xyz, rl, return_info=True)
    assert np.array_equal(got, want) and ginfo["D"] == info["D"] and ginfo["votes"] == info["votes"]
    # empty mask inside a batch is a per-item status, not an error
    from rcvpose_b200 import api
    depth = torch.zeros((1, 480, 640), dtype=torch.int16, device="cuda")
    radius = torch.ones((1, 1, 480, 640), dtype=torch.float32, device="cuda")
    out = ctx.vote_frames(depth, radius, torch.from_numpy(synth.linemod_K).cuda(), mask_flags=api.RCV_MASK_RADIUS_NONZERO)
    assert int(out["status"][0, 0]) == api.RCV_ST_EMPTY_MASK


def test_ragged_batch_capacity_and_error_codes():
    """The boundary's error behaviour (include/rcvvote.h): data-dependent conditions are per-item status bits and never UB
    (empty item inside a batch, D beyond max_grid, point pool or tile work list exhausted: the other items of the call
    are still bit-exact); bad arguments and requests beyond the capacities fixed at rcv_create are call errors."""
    from rcvpose_b200 import api
    rng = np.random.default_rng(21)
    sizes = [0, 5, 300, 1, 120]
    clouds, radii = [], []
    for n in sizes:
        xyz = rng.normal(0, 0.02, size=(n, 3)) + np.array([0.0, 0.05, 0.7])
        kp = np.array([0.07, 0.02, 0.66])
        clouds.append(xyz)
        radii.append((np.linalg.norm(xyz - kp, axis=1) * 10).astype(np.float32))
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64).cuda()
    X = torch.from_numpy(np.concatenate(clouds)).cuda()
    Rr = torch.from_numpy(np.concatenate(radii)).cuda()
    want = [oracle.Accumulator_3D(c, r, return_info=True) if len(c) else None for c, r in zip(clouds, radii)]

    def check(out, ok_items):
        for b in ok_items:
            w, info = want[b]
            assert int(out["status"][b]) == 0 and int(out["grid"][b]) == info["D"] and int(out["votes"][b]) == info["votes"]
            assert np.array_equal(out["centre_mm"][b].cpu().numpy(), w[0])

    ctx = api.VoteContext(0, max_items=8, max_points_total=4096, max_grid=128)
    out = ctx.vote_points(X, Rr, off)                                   # ragged batch with an empty item
    torch.cuda.synchronize()
    assert int(out["status"][0]) == api.RCV_ST_EMPTY_MASK
    check(out, [1, 2, 3, 4])
    # grid larger than the capacity of the context: that item only
    small = api.VoteContext(0, max_items=8, max_points_total=4096, max_grid=max(want[1][1]["D"], want[3][1]["D"]))
    out = small.vote_points(X, Rr, off)
    torch.cuda.synchronize()
    big = [b for b in (1, 2, 3, 4) if want[b][1]["D"] > small.max_grid]
    assert big and all(int(out["status"][b]) & api.RCV_ST_D_EXCEEDS_CAP for b in big)
    check(out, [b for b in (1, 2, 3, 4) if b not in big])
    # tile work list too short: items that do not fit report it, earlier ones are unaffected
    few = api.VoteContext(0, max_items=8, max_points_total=4096, max_grid=128, max_units=3)
    out = few.vote_points(X, Rr, off)
    torch.cuda.synchronize()
    st = [int(v) for v in out["status"].cpu()]
    assert any(v & api.RCV_ST_UNIT_OVERFLOW for v in st)
    check(out, [b for b in (1, 2, 3, 4) if st[b] == 0])
    # point pool exhausted by the frames entry point: overflowed items carry the status and no points
    frames = [synth.config3_frame(f) for f in range(2)]
    depth = torch.from_numpy(np.stack([f["depth"] for f in frames]).view(np.int16)).cuda()
    radius = torch.from_numpy(np.stack([f["radius"] for f in frames])).cuda()
    K = torch.from_numpy(synth.linemod_K).cuda()
    full = api.VoteContext(0, max_items=8, max_points_total=1 << 16, max_grid=256).vote_frames(depth, radius, K, mask_flags=api.RCV_MASK_RADIUS_NONZERO)
    n0 = int(full["n_points"][0, 0])
    tight = api.VoteContext(0, max_items=8, max_points_total=n0 + 10, max_grid=256)
    out = tight.vote_frames(depth, radius, K, mask_flags=api.RCV_MASK_RADIUS_NONZERO)
    torch.cuda.synchronize()
    assert int(out["status"][0, 0]) == 0 and np.array_equal(out["centre_mm"][0, 0].cpu().numpy(), full["centre_mm"][0, 0].cpu().numpy())
    assert all(int(v) & api.RCV_ST_POINT_OVERFLOW for v in out["status"].flatten()[1:].cpu())
    assert int(out["n_points"].flatten()[1:].sum()) == 0
    # call errors
    with pytest.raises(api.RcvError):
        tight.vote_frames(depth.repeat(5, 1, 1), radius.repeat(5, 1, 1, 1), K, mask_flags=api.RCV_MASK_RADIUS_NONZERO)   # 30 items > max_items
    with pytest.raises(api.RcvError):
        ctx.vote_points(X, Rr, off, acc_unit=0.0)
    with pytest.raises(api.RcvError):
        ctx.vote_frames(depth, radius, K, mask_flags=api.RCV_MASK_SEM_GT)                                                   # sem rule without a sem map
    out = ctx.vote_points(X, Rr, off)                                   # the context is still usable after errors
    torch.cuda.synchronize()
    check(out, [1, 2, 3, 4])


def test_argmax_volume_first_max_in_c_order(ctx):
    rng = np.random.default_rng(2)
    for D in (5, 33, 86):
        v = rng.integers(0, 50, size=(D, D, D)).astype(np.int32)
        v[rng.integers(0, D), rng.integers(0, D), rng.integers(0, D)] = 77
        v.ravel()[D * D * D // 2 + 3] = 77                      # a tie: the smaller linear index must win
        idx, mx = ctx.argmax_volume(torch.from_numpy(v).cuda())
        want = np.argwhere(v == v.max())[0]
        assert np.array_equal(idx.cpu().numpy(), want) and int(mx) == 77
    z = torch.zeros((7, 7, 7), dtype=torch.int32, device="cuda")
    idx, mx = ctx.argmax_volume(z)
    assert idx.tolist() == [0, 0, 0] and int(mx) == 0


@pytest.mark.parametrize("t", range(6))
def test_horn_vs_reference(golden, t):
    from rcvpose_b200.util.horn import HornPoseFitting
    P1, P2, RT = golden["h%d_P1" % t], golden["h%d_P2" % t], golden["h%d_RT" % t]
    A = np.zeros((4, 4))
    a, b = P1.copy(), P2.copy()
    assert HornPoseFitting().lmshorn(a, b, P1.shape[0], A) is None
    assert np.array_equal(a, P1) and np.array_equal(b, P2)
    np.testing.assert_allclose(A[:3, :3], RT[:3, :3], rtol=0, atol=1e-12)
    np.testing.assert_allclose(A[:3, 3], RT[:3, 3], rtol=1e-12, atol=1e-9)
    assert np.array_equal(A[3], [0, 0, 0, 1])


def test_horn_batch_device(ctx, golden):
    P1 = torch.from_numpy(golden["h0_P1"]).cuda()
    est = torch.from_numpy(np.stack([golden["h%d_P2" % t] for t in range(4)])).cuda()
    mdl = torch.from_numpy(np.stack([golden["h%d_P1" % t] for t in range(4)])).cuda()
    RT = ctx.horn_batch(mdl, est).cpu().numpy()
    for t in range(4):
        np.testing.assert_allclose(RT[t], golden["h%d_RT" % t], rtol=1e-12, atol=1e-9)
    RT0 = ctx.horn_batch(P1, est[:1]).cpu().numpy()
    np.testing.assert_allclose(RT0[0], golden["h0_RT"], rtol=1e-12, atol=1e-9)


def test_add_metric_vs_kdtree(ctx):
    """ADD(-S) before ICP (AccumulatorSpace.py:664-702): brute-force float64 nearest neighbour on the GPU against a k-d tree
    on the host; 1e-9 relative (the nearest neighbour is exact on both sides, only the summation order differs)."""
    rng = np.random.default_rng(31)
    M, B = 1357, 6
    u = rng.normal(size=(M, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    model = u * np.array([45.0, 30.0, 60.0]) * rng.uniform(0.7, 1.0, size=(M, 1))     # an ellipsoidal shell of CAD points, mm

    def pose(rv, t):
        th = np.linalg.norm(rv); k = rv / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
        RT = np.eye(4); RT[:3, :3] = R; RT[:3, 3] = t
        return RT
    RT_gt = np.stack([pose(rng.normal(size=3), rng.uniform(-200, 200, 3) + np.array([0, 0, 900.0])) for _ in range(B)])
    RT_est = RT_gt.copy()
    for b in range(B):                                   # frame 0: identical pose (distance 0); then growing errors
        if b:
            RT_est[b] = pose(rng.normal(size=3) * 0.01 * b, rng.normal(size=3) * 2.0 * b) @ RT_gt[b]
    mean, mn = ctx.add_metric(torch.from_numpy(model).cuda(), torch.from_numpy(RT_est).cuda(), torch.from_numpy(RT_gt).cuda())
    torch.cuda.synchronize()
    for b in range(B):
        wm, wn = oracle.add_metric(model, RT_est[b], RT_gt[b])
        assert abs(float(mean[b]) - wm) <= 1e-9 * max(1.0, wm), (b, float(mean[b]), wm)
        assert abs(float(mn[b]) - wn) <= 1e-9 * max(1.0, wn), (b, float(mn[b]), wn)
    assert float(mean[0]) == 0.0 and float(mean[B - 1]) > float(mean[1]) > 0.0


def test_add_metric_model_grid_equals_all_pairs(ctx, monkeypatch):
    """ADD(-S) through the grid over the CAD model (refine.cu, default for models of 1024 points and more) against the all-pairs
    kernel (RCV_ADD_BRUTE=1): the same squared distances and the same reduction tree, so mean and minimum are bit-identical --
    for nearly equal poses (one or two shells of cells), a grossly wrong pose (the search visits the whole grid), a pose that
    puts the query cloud outside the grid's box, and a flat model (a grid one cell thick)."""
    rng = np.random.default_rng(33)

    def pose(rv, t):
        th = np.linalg.norm(rv); k = rv / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        RT = np.eye(4); RT[:3, :3] = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx; RT[:3, 3] = t
        return RT
    for M, flat in ((5841, False), (2048, True)):
        u = rng.normal(size=(M, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
        model = u * np.array([45.0, 30.0, 60.0]) * rng.uniform(0.8, 1.0, size=(M, 1))
        if flat:
            model[:, 2] = 0.0
        B = 7
        RT_gt = np.stack([pose(rng.normal(size=3), rng.uniform(-200, 200, 3) + np.array([0, 0, 900.0])) for _ in range(B)])
        RT_est = RT_gt.copy()
        for b in range(1, B):
            RT_est[b] = pose(rng.normal(size=3) * 0.01 * b, rng.normal(size=3) * 1.5 * b) @ RT_gt[b]
        RT_est[5] = pose(rng.normal(size=3) * 2.0, np.array([40.0, -25.0, 30.0])) @ RT_gt[5]          # grossly wrong rotation
        RT_est[6] = pose(rng.normal(size=3) * 0.01, np.array([500.0, 300.0, -400.0])) @ RT_gt[6]       # half a metre away
        args = (torch.from_numpy(model).cuda(), torch.from_numpy(RT_est).cuda(), torch.from_numpy(RT_gt).cuda())
        monkeypatch.setenv("RCV_ADD_BRUTE", "1")
        want = [t.cpu().numpy() for t in ctx.add_metric(*args)]
        monkeypatch.setenv("RCV_ADD_BRUTE", "0")
        got = [t.cpu().numpy() for t in ctx.add_metric(*args)]
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (M, got, want)
        assert got[0][0] == 0.0 and got[0][6] > 100.0 and np.isfinite(got[0]).all()


@pytest.mark.parametrize("shape", [(1, 480, 640), (3, 480, 640), (2, 24, 40), (1, 8, 24)])
def test_head_1x1_tensor_core_vs_torch(ctx, shape):
    """K5 (conv8 of the producer, models/fcnresnet.py:118,187-189): tcgen05 kernel vs torch's fp32 conv of the same
    bf16 operands.  Products of bf16 values are exact in fp32; only the accumulation order differs: 1e-5 relative."""
    B, H, W = shape
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H)
    up = torch.relu(torch.randn((B, 32, H, W), generator=g, device="cuda")).to(torch.bfloat16)     # post BN+ReLU activations
    weight = torch.randn((2, 32, 1, 1), generator=g, device="cuda") * 0.2
    bias = torch.randn((2,), generator=g, device="cuda")
    got = ctx.head_1x1(up, weight, bias)
    torch.cuda.synchronize()
    want = torch.nn.functional.conv2d(up.float(), weight.to(torch.bfloat16).float(), bias)
    assert got.shape == (B, 2, H, W) and got.dtype == torch.float32
    err = (got - want).abs().max().item()
    scale = want.abs().max().item()
    assert err <= 1e-5 * scale + 1e-6, (err, scale)
    # seg / radius planes feed the vote directly: a second call on the same stream gives the same bits
    assert torch.equal(got, ctx.head_1x1(up, weight, bias))
