"""Fuzzes the rasteriser arithmetic (rcvpose_b200/csrc/raster_core.h, compiled for the host by
tests/hostsim.cpp) against the oracle's brute-force restatement of fast_for
(AccumulatorSpace.py:325-341).  The vote volume must be bit-exact: integer adds commute."""
import numpy as np
import pytest

from oracle import oracle
from tests import hostsim

RCV_MIN_R = 1


def _check(p, R, D, **kw):
    want = oracle.fast_for(p, R, D)
    got, st = hostsim.render(p, R, D, **kw)
    bad = np.argwhere(got != want)
    assert bad.size == 0, "first mismatch at %s: got %d want %d (of %d)" % (bad[0], got[tuple(bad[0])], want[tuple(bad[0])], len(bad))
    assert st["votes"] == int(want.sum())
    return st


@pytest.mark.parametrize("seed", range(6))
def test_random_points_all_radii(seed):
    rng = np.random.default_rng(seed)
    D = int(rng.integers(20, 70))
    n = 40
    p = rng.uniform(-3, D + 3, size=(n, 3))
    R = rng.integers(-1, D // 2 + 4, size=n).astype(np.int32)
    _check(p, R, D, sqrt_perturb=seed % 2)


@pytest.mark.parametrize("R", [1, 2, 3, 4, 5, 6, 7, 8, 10, 13, 17, 24, 31, 40])
def test_single_radius_many_offsets(R):
    rng = np.random.default_rng(100 + R)
    D = 2 * R + 9
    p = (D / 2.0) + rng.uniform(-1.0, 1.0, size=(24, 3))
    _check(p, np.full(24, R, np.int32), D, sqrt_perturb=1)


def test_lattice_aligned_and_half_integer_points():
    # adversarial: exact boundary hits (|v-p| == R exactly), half-integer ties, zero fractions
    D = 41
    pts, Rs = [], []
    for R in (1, 2, 3, 5, 10, 13, 15):  # 5,10,13,15: many integer triples with x^2+y^2+z^2 == R^2
        for off in ((0, 0, 0), (0.5, 0, 0), (0.5, 0.5, 0.5), (0.25, -0.25, 0.5), (1e-9, 0, 0), (0, -1e-12, 0.5 - 1e-12)):
            pts.append(np.array([20.0, 20.0, 20.0]) + np.array(off))
            Rs.append(R)
    p, R = np.array(pts), np.array(Rs, np.int32)
    _check(p, R, D)
    _check(p, R, D, sqrt_perturb=1)


def test_clipping_at_grid_faces_and_corners():
    rng = np.random.default_rng(5)
    D = 30
    p = np.concatenate([rng.uniform(-2.5, 2.5, size=(10, 3)), D - 1 + rng.uniform(-2.5, 2.5, size=(10, 3)),
                        np.array([[0.0, 15.2, 29.0], [-2.4, -2.4, -2.4], [31.9, 31.9, 31.9]])])
    R = rng.integers(1, 20, size=p.shape[0]).astype(np.int32)
    _check(p, R, D)


@pytest.mark.parametrize("tile", [(0, 5, 0, 37), (5, 7, 0, 37), (30, 7, 0, 37), (11, 1, 0, 37), (11, 1, 8, 13), (3, 2, 30, 7), (0, 37, 0, 37)])
def test_tiles_partition_the_volume(tile):
    rng = np.random.default_rng(9)
    D = 37
    p = rng.uniform(5, 32, size=(30, 3))
    R = rng.integers(1, 16, size=30).astype(np.int32)
    want = oracle.fast_for(p, R, D)
    i0, ni, j0, nj = tile
    got, _ = hostsim.render(p, R, D, tile=tile, Dp=39)
    assert np.array_equal(got, want[i0:i0 + ni, j0:j0 + nj, :])


def test_large_radius_thick_rings():
    rng = np.random.default_rng(21)
    D = 120
    p = rng.uniform(40, 80, size=(6, 3))
    R = np.array([37, 45, 52, 58, 29, 33], np.int32)
    want = oracle.fast_for(p, R, D, method="scatter")
    got, st = hostsim.render(p, R, D, sqrt_perturb=1)
    assert np.array_equal(got, want)


def test_linemod_shaped_frame_matches_reference_volume(golden):
    from rcvpose_b200 import synth
    fr = synth.config1_frame()
    xyz, rl = synth.frame_to_points(fr["K"], fr["depth"], fr["radius"][0])
    pre = oracle.prelude(xyz, rl)
    got, st = hostsim.render(pre["p"], pre["R"], pre["D"])
    assert np.array_equal(got, golden["c1_volume"])   # the real reference's volume, every voxel
    assert st["votes"] == int(golden["c1_votes"])


@pytest.mark.parametrize("slab", [1, 2, 3, 4, 6, 7, 12, 32])
def test_slab_thickness_and_chunking(slab):
    """The kernel walks a slab in chunks of 3 or 4 slices and draws the polar caps once per slab: every slab
    thickness must give the same volume (exercises chunk alignment, polar masks and the annulus row ranges)."""
    rng = np.random.default_rng(40 + slab)
    D = 64
    n = 70
    p = rng.uniform(14, 50, size=(n, 3))
    R = rng.integers(5, 30, size=n).astype(np.int32)
    want = oracle.fast_for(p, R, D, method="scatter")
    got, st = hostsim.render(p, R, D, slab=slab, sqrt_perturb=slab % 2)
    assert np.array_equal(got, want)
    assert st["polar_cells"] > 0 and st["lane_tasks"] > 0


@pytest.mark.parametrize("R", [6, 7, 8, 9, 11, 12, 14, 19])
def test_polar_threshold_radii(R):
    """R = RCV_POLAR_MIN_R - 1 (dense scan), R = RCV_POLAR_MIN_R and just above (polar pass, one voxel per column)."""
    rng = np.random.default_rng(300 + R)
    D = 2 * R + 12
    p = (D / 2.0) + rng.uniform(-2.0, 2.0, size=(64, 3))
    for slab in (5, 32):
        want = oracle.fast_for(p, np.full(64, R, np.int32), D)
        got, _ = hostsim.render(p, np.full(64, R, np.int32), D, slab=slab, sqrt_perturb=1)
        assert np.array_equal(got, want)


def test_surface_patch_like_a_frame_row():
    """Points of one warp like consecutive pixels of an image row (close in y, spread in x, varying R)."""
    rng = np.random.default_rng(77)
    D = 90
    n = 96
    x = np.linspace(30, 50, n) + rng.normal(0, 0.05, n)
    p = np.stack([x, 40 + rng.normal(0, 0.3, n), 45 + 0.02 * (x - 40) ** 2], axis=1)
    R = np.round(np.linalg.norm(p - np.array([60.0, 55.0, 30.0]), axis=1)).astype(np.int32)
    want = oracle.fast_for(p, R, D, method="scatter")
    for slab in (6, 9):
        got, _ = hostsim.render(p, R, D, slab=slab)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("seed", range(8))
def test_fast_ring_pass_interior_spheres(seed):
    """Spheres that cannot leave the tile take the fast ring pass (ring2: magic-number addresses, interior columns
    without the ownership test); groups of 32 like a warp, mixed radii, device-like sqrt error."""
    rng = np.random.default_rng(700 + seed)
    Rmax = int(rng.integers(8, 30))
    D = 2 * Rmax + 12
    n = 64
    R = rng.integers(RCV_MIN_R, Rmax + 1, size=n).astype(np.int32)
    lo = R[:, None] + 3.0
    p = lo + rng.uniform(0, 1, size=(n, 3)) * (D - 1 - 2 * lo)
    if seed % 2:                                  # a warp that shares (y voxel, R) like the kernel's vote order
        R[:] = Rmax
        p = (D / 2.0) + rng.uniform(-1.0, 1.0, size=(n, 3))
    st = _check(p, R, D, sqrt_perturb=seed % 3 != 0)
    assert st["ring2_tasks"] > 0


def test_fast_ring_pass_lattice_ties():
    # lattice-aligned and half-integer points: exact ties of the magic-number rounding and exact boundary hits
    D = 64
    pts, Rs = [], []
    for R in (7, 10, 13, 15, 17, 25):
        for off in ((0, 0, 0), (0.5, 0, 0), (0.5, 0.5, 0.5), (0.25, -0.25, 0.5), (1e-9, 0, 0), (0, -1e-12, 0.5 - 1e-12), (-0.5, 0.5, 0)):
            pts.append(np.array([31.0, 32.0, 31.0]) + np.array(off))
            Rs.append(R)
    while len(pts) % 32:
        pts.append(np.array([31.5, 31.25, 32.0])); Rs.append(12)
    p, R = np.array(pts), np.array(Rs, np.int32)
    for perturb in (0, 1):
        st = _check(p, R, D, sqrt_perturb=perturb)
        assert st["ring2_tasks"] > 0
