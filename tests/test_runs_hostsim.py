"""Fuzzes the run-length rasteriser arithmetic (rcvpose_b200/csrc/runs_core.h, compiled for the host by
tests/hostsim_runs.cpp) against the oracle's brute-force restatement of fast_for (AccumulatorSpace.py:325-341).
The vote volume must be bit-exact: the difference array holds integer adds, which commute."""
import numpy as np
import pytest

from oracle import oracle
from tests import hostsim


def _check(p, R, D, method="brute", **kw):
    want = oracle.fast_for(p, R, D, method=method)
    got, st = hostsim.render_runs(p, R, D, **kw)
    bad = np.argwhere(got != want)
    assert bad.size == 0, "first mismatch at %s: got %d want %d (of %d)" % (bad[0], got[tuple(bad[0])], want[tuple(bad[0])], len(bad))
    return st


def _guards(p, R, D):
    """Guard cells the kernel's prelude would give this item: boundaries lie within R + 2 of the nearest lattice point."""
    live = R > 0
    if not live.any():
        return 0, 0
    lo = (p[live].min(axis=1) - R[live]).min()
    hi = (p[live].max(axis=1) + R[live]).max()
    return max(0, int(np.ceil(2.5 - lo))), max(0, int(np.ceil(hi + 2.5 - (D - 1))))


@pytest.mark.parametrize("seed", range(8))
def test_random_points_all_radii(seed):
    rng = np.random.default_rng(seed)
    D = int(rng.integers(20, 70))
    n = 40
    p = rng.uniform(2, D - 2, size=(n, 3))
    R = rng.integers(-1, D // 2 + 4, size=n).astype(np.int32)
    glo, ghi = _guards(p, R, D)
    clip = glo > 8 or ghi > 8
    if clip:
        glo, ghi = 0, 0
    _check(p, R, D, glo=glo, ghi=ghi, clip=clip, slab=int(rng.integers(1, 9)), sqrt_perturb=seed % 2)


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 6, 7, 8, 10, 13, 17, 24, 31, 40])
def test_single_radius_many_offsets(R):
    rng = np.random.default_rng(100 + R)
    D = 2 * R + 9
    p = (D / 2.0) + rng.uniform(-1.0, 1.0, size=(24, 3))
    _check(p, np.full(24, R, np.int32), D, sqrt_perturb=1)


def test_lattice_aligned_and_half_integer_points():
    # adversarial: exact boundary hits (|v-p| == R exactly), half-integer ties of the rounding, zero fractions
    D = 41
    pts, Rs = [], []
    for R in (1, 2, 3, 5, 10, 13, 15):  # 5,10,13,15: many integer triples with x^2+y^2+z^2 == R^2
        for off in ((0, 0, 0), (0.5, 0, 0), (0, 0, 0.5), (0.5, 0.5, 0.5), (0.25, -0.25, 0.5), (1e-9, 0, 0), (0, 0, 1e-9), (0, -1e-12, 0.5 - 1e-12),
                    (0, 0, -0.5), (1e-5, 1e-5, 1e-5), (0.0, 0.0, 0.02)):
            pts.append(np.array([20.0, 20.0, 20.0]) + np.array(off))
            Rs.append(R)
    p, R = np.array(pts), np.array(Rs, np.int32)
    for perturb in (0, 1):
        for slab in (1, 4):
            st = _check(p, R, D, sqrt_perturb=perturb, slab=slab)
            assert st["exact_calls"] > 0


def test_tangent_columns_and_tiny_discs():
    """Columns that graze the sphere (g ~ 0) and spheres whose cap in a slice is smaller than a voxel: the float32
    transition is meaningless there and everything rests on the flag + exact path."""
    rng = np.random.default_rng(11)
    D = 48
    pts, Rs = [], []
    for R in (3, 6, 12, 20):
        for _ in range(12):
            base = np.array([24.0, 24.0, 24.0]) + rng.integers(-2, 3, size=3)
            d = rng.normal(size=3); d /= np.linalg.norm(d)
            k = rng.integers(0, 3)
            off = np.zeros(3); off[k] = R - rng.choice([0.0, 1e-7, 1e-4, 1e-3, 0.02, 0.43, 0.4330127018922193])   # voxel exactly / nearly on the shell along an axis
            q = base - off
            pts.append(q + rng.choice([0.0, 1e-9, 1e-6]) * d)
            Rs.append(R)
    p, R = np.array(pts), np.array(Rs, np.int32)
    for perturb in (0, 1):
        _check(p, R, D, glo=0, ghi=0, clip=True, sqrt_perturb=perturb, slab=3)


def test_clipping_at_grid_faces_and_corners():
    rng = np.random.default_rng(5)
    D = 30
    p = np.concatenate([rng.uniform(-2.5, 2.5, size=(10, 3)), D - 1 + rng.uniform(-2.5, 2.5, size=(10, 3)),
                        np.array([[0.0, 15.2, 29.0], [-2.4, -2.4, -2.4], [31.9, 31.9, 31.9]])])
    R = rng.integers(1, 20, size=p.shape[0]).astype(np.int32)
    _check(p, R, D, glo=0, ghi=0, clip=True)             # no guard at all: every boundary is clamped
    _check(p, R, D, glo=3, ghi=2, clip=True, slab=5)


def test_guard_band_holds_overhanging_spheres():
    """The reference sizes the grid so that spheres overhang it by a few voxels; with guard cells the unclipped variant draws them."""
    rng = np.random.default_rng(6)
    D = 40
    R = rng.integers(3, 12, size=64).astype(np.int32)
    p = np.where(rng.random((64, 3)) < 0.5, R[:, None] - 2.4 + rng.random((64, 3)), D - 1 - R[:, None] + 2.4 - rng.random((64, 3)))
    glo, ghi = _guards(p, R, D)
    assert 0 < glo <= 8 and 0 < ghi <= 8
    _check(p, R, D, glo=glo, ghi=ghi, slab=4)
    # guard rows as well as guard cells (what a tile looked like before the rows were restricted to the grid's): same counts
    got, _ = hostsim.render_runs(p, R, D, glo=glo, ghi=ghi, slab=4, band=(-glo, D + glo + ghi))
    assert np.array_equal(got, oracle.fast_for(p, R, D))


@pytest.mark.parametrize("band", [(0, 37), (0, 12), (12, 13), (25, 12), (8, 1), (30, 7)])
def test_row_bands_partition_the_volume(band):
    rng = np.random.default_rng(9)
    D = 37
    p = rng.uniform(5, 32, size=(30, 3))
    R = rng.integers(1, 16, size=30).astype(np.int32)
    want = oracle.fast_for(p, R, D)
    j0, nj = band
    got, _ = hostsim.render_runs(p, R, D, glo=0, ghi=0, band=band, clip=True, slab=3)
    assert np.array_equal(got[j0:j0 + nj], want[j0:j0 + nj])
    assert not got[:j0].any() and not got[j0 + nj:].any()


@pytest.mark.parametrize("band", [(0, 12), (12, 13), (25, 12), (8, 1), (30, 7), (0, 37)])
@pytest.mark.parametrize("NC", [1, 2])
def test_guarded_row_bands_need_no_clipping(band, NC):
    """Row bands of a large grid keep the guard cells along z: the rows are restricted by the lanes' column ranges, the cells
    by the guard, so the unclipped column loops draw a band (spheres overhang the grid on both sides here)."""
    rng = np.random.default_rng(10)
    D = 37
    R = rng.integers(2, 14, size=40).astype(np.int32)
    p = R[:, None] + 1.0 + rng.random((40, 3)) * (D - 3.0 - 2.0 * R[:, None])
    p[:8] = np.where(rng.random((8, 3)) < 0.5, R[:8, None] - 2.2 + rng.random((8, 3)), D - 1 - R[:8, None] + 2.2 - rng.random((8, 3)))
    glo, ghi = _guards(p, R, D)
    assert 0 < glo <= 8 and 0 < ghi <= 8
    want = oracle.fast_for(p, R, D)
    j0, nj = band
    got, _ = hostsim.render_runs(p, R, D, glo=glo, ghi=ghi, band=band, clip=False, slab=1, NC=NC, sqrt_perturb=1)
    assert np.array_equal(got[j0:j0 + nj], want[j0:j0 + nj])
    assert not got[:j0].any() and not got[j0 + nj:].any()


@pytest.mark.parametrize("seed", range(6))
def test_tiles_cover_only_the_box_the_spheres_reach(seed):
    """The prelude's windows: along z a tile stores the cells [min(z - R) - 2.5, max(z + R) + 2.5] only -- SIGNED guards, negative
    where the spheres stay inside the grid -- and along x the rows the spheres can reach.  The reference sizes the grid by
    the widest axis, so the other axes have slack (here: a cloud that is flat along z and off-centre along x)."""
    rng = np.random.default_rng(100 + seed)
    D = 64
    n = 80
    R = rng.integers(1, 12, size=n).astype(np.int32)
    p = np.stack([rng.uniform(14, 40, n), rng.uniform(13, 51, n), rng.uniform(28, 36, n)], axis=1)
    if seed % 2:
        p[:, 2] += 14.0                                    # the window touches / overhangs the top of the grid instead
    lo = (p - R[:, None]).min(axis=0) - 2.5
    hi = (p + R[:, None]).max(axis=0) + 2.5
    glo, ghi = -int(np.floor(lo[2])), int(np.ceil(hi[2])) - (D - 1)
    assert glo < 0 and (ghi < 0 or seed % 2)
    x0, x1 = max(0, int(np.floor(lo[0]))), min(D - 1, int(np.ceil(hi[0])))
    want = oracle.fast_for(p, R, D)
    for slab, NC in ((3, 3), (5, 2), (4, 4)):
        got, _ = hostsim.render_runs(p, R, D, glo=glo, ghi=min(ghi, 8), band=(x0, x1 - x0 + 1), slab=slab, NC=NC, sqrt_perturb=seed % 2)
        assert np.array_equal(got, want), (slab, NC)


def test_large_radius_thick_rings():
    rng = np.random.default_rng(21)
    D = 120
    p = rng.uniform(40, 80, size=(6, 3))
    R = np.array([37, 45, 52, 58, 29, 33], np.int32)
    glo, ghi = _guards(p, R, D)
    _check(p, R, D, method="scatter", glo=min(glo, 8), ghi=min(ghi, 8), clip=(glo > 8 or ghi > 8), sqrt_perturb=1)


def test_very_large_radius():
    rng = np.random.default_rng(22)
    D = 400
    p = rng.uniform(180, 220, size=(3, 3))
    R = np.array([150, 171, 96], np.int32)
    want = oracle.fast_for(p, R, D, method="scatter")
    got, st = hostsim.render_runs(p, R, D, glo=0, ghi=0, clip=True, slab=25, NC=1, sqrt_perturb=1)
    assert np.array_equal(got, want)


def test_linemod_shaped_frame_matches_reference_volume(golden):
    from rcvpose_b200 import synth
    fr = synth.config1_frame()
    xyz, rl = synth.frame_to_points(fr["K"], fr["depth"], fr["radius"][0])
    pre = oracle.prelude(xyz, rl)
    glo, ghi = _guards(pre["p"], pre["R"], pre["D"])
    got, st = hostsim.render_runs(pre["p"], pre["R"], pre["D"], glo=glo, ghi=ghi)
    assert np.array_equal(got, golden["c1_volume"])   # the real reference's volume, every voxel
    assert int(got.sum()) == int(golden["c1_votes"])
    assert st["exact_calls"] < 0.002 * st["atomics"]


@pytest.mark.parametrize("slab", [1, 2, 3, 4, 6, 7, 12, 32])
def test_slab_thickness_and_chunking(slab):
    rng = np.random.default_rng(40 + slab)
    D = 64
    n = 70
    p = rng.uniform(14, 50, size=(n, 3))
    R = rng.integers(5, 30, size=n).astype(np.int32)
    glo, ghi = _guards(p, R, D)
    for NC in sorted({1, min(slab, 2), min(slab, 3), min(slab, 4)}):
        _check(p, R, D, method="scatter", glo=glo, ghi=ghi, slab=slab, NC=NC, sqrt_perturb=slab % 2)


def test_surface_patch_like_a_frame_row():
    """Points of one warp like consecutive pixels of an image row (close in y, spread in x, varying R)."""
    rng = np.random.default_rng(77)
    D = 90
    n = 96
    x = np.linspace(30, 50, n) + rng.normal(0, 0.05, n)
    p = np.stack([x, 40 + rng.normal(0, 0.3, n), 45 + 0.02 * (x - 40) ** 2], axis=1)
    R = np.round(np.linalg.norm(p - np.array([60.0, 55.0, 30.0]), axis=1)).astype(np.int32)
    glo, ghi = _guards(p, R, D)
    for slab in (6, 9):
        _check(p, R, D, method="scatter", glo=glo, ghi=ghi, slab=slab)


@pytest.mark.parametrize("seed", range(6))
def test_mixed_radii_in_one_warp_park_their_columns(seed):
    """Lanes with a small sphere next to lanes with a large one: the loop runs to the warp maximum and the small lanes'
    surplus columns must stay inside the tile (parked on the lane's own edge row) and draw nothing."""
    rng = np.random.default_rng(700 + seed)
    D = 70
    n = 64
    R = np.where(rng.random(n) < 0.5, rng.integers(1, 4, size=n), rng.integers(20, 30, size=n)).astype(np.int32)
    p = np.where((R < 5)[:, None], rng.uniform(0.5, D - 1.5, size=(n, 3)), rng.uniform(28, 40, size=(n, 3)))
    glo, ghi = _guards(p, R, D)
    _check(p, R, D, method="scatter", glo=glo, ghi=ghi, slab=4, sqrt_perturb=seed % 2)
