"""Long-running fuzz of the run-length rasteriser (host build) against the oracle.  usage: fuzz_runs.py [seconds] [seed0]"""
import sys, time
import numpy as np
sys.path.insert(0, __file__.rsplit("/", 2)[0])
from oracle import oracle
from tests import hostsim

T = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
t0 = time.time(); cases = 0; votes = 0; flagged = 0; fixes = 0; atomics = 0
while time.time() - t0 < T:
    rng = np.random.default_rng(seed); seed += 1
    kind = seed % 4
    if kind == 0:       # near-lattice points, medium radii
        D = int(rng.integers(40, 100)); n = 48
        R = rng.integers(1, D // 2 - 2, size=n).astype(np.int32)
        p = np.round(rng.uniform(R[:, None] + 1, D - 2 - R[:, None], size=(n, 3))) + rng.choice([0, 0.5, 1e-7, -1e-7, 0.25, 1e-3], size=(n, 3)) * rng.choice([1, -1], size=(n, 3))
    elif kind == 1:     # surface patch like a frame (sorted by y voxel, R)
        D = int(rng.integers(60, 130)); n = 160
        c0 = rng.uniform(D * 0.35, D * 0.65, size=3); k = c0 + rng.normal(size=3) * D * 0.15
        d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        p = c0 + d * rng.uniform(5, 10)
        R = np.round(np.linalg.norm(p - k, axis=1) + rng.normal(0, 0.2, n)).astype(np.int32)
        o = np.lexsort((R, np.rint(p[:, 1]))); p, R = p[o], R[o]
    elif kind == 2:     # big radii
        D = int(rng.integers(150, 260)); n = 8
        R = rng.integers(40, D // 2 - 5, size=n).astype(np.int32)
        p = rng.uniform(R[:, None] + 3, D - 4 - R[:, None], size=(n, 3))
    else:               # anything, clipped
        D = int(rng.integers(10, 60)); n = 64
        R = rng.integers(-1, D, size=n).astype(np.int32)
        p = rng.uniform(-3, D + 3, size=(n, 3))
    live = R > 0
    lo = (p[live].min(axis=1) - R[live]).min() if live.any() else 0; hi = (p[live].max(axis=1) + R[live]).max() if live.any() else 0
    glo, ghi = max(0, int(np.ceil(2.5 - lo))), max(0, int(np.ceil(hi + 2.5 - (D - 1))))
    clip = glo > 8 or ghi > 8
    if clip: glo = ghi = 0
    slab = int(rng.integers(1, 9))
    want = oracle.fast_for(p, R, D, method="scatter" if D > 64 else "brute")
    NC = min(slab, int(rng.integers(1, 5)))
    window = None
    if not clip and live.any() and seed % 5 in (1, 2):   # the prelude's windows: signed guards along z, reachable rows along x
        wlo = (p[live] - R[live, None]).min(axis=0) - 2.5; whi = (p[live] + R[live, None]).max(axis=0) + 2.5
        g0, g1 = -int(np.floor(wlo[2])), int(np.ceil(whi[2])) - (D - 1)
        if g0 > -D and g1 > -D and g0 <= 8 and g1 <= 8 and D + g0 + g1 >= 1:
            glo, ghi = g0, g1
        x0, x1 = max(0, int(np.floor(wlo[0]))), min(D - 1, int(np.ceil(whi[0])))
        window = (x0, x1 - x0 + 1) if x0 <= x1 else (0, 1)
    if seed % 3 == 0:   # row bands (large-grid tiles): the rows of a band are restricted by the lanes' column ranges, guards along z only
        nj = int(rng.integers(1, max(2, D // 2)))
        got = np.zeros((D, D, D), np.int32); st = dict(flagged_cols=0, fixes=0, atomics=0)
        r0, rn = window if window is not None else (0, D)
        for j0 in range(r0, r0 + rn, nj):
            g1, s1 = hostsim.render_runs(p, R, D, glo=glo, ghi=ghi, band=(j0, min(nj, r0 + rn - j0)), clip=clip, slab=slab, NC=NC, sqrt_perturb=seed % 2)
            got += g1
            for k in st: st[k] += s1[k]
    else:
        got, st = hostsim.render_runs(p, R, D, glo=glo, ghi=ghi, band=window, clip=clip, slab=slab, NC=NC, sqrt_perturb=seed % 2)
    if not np.array_equal(got, want):
        print("MISMATCH seed", seed - 1, "kind", kind, "ndiff", int((got != want).sum())); sys.exit(1)
    cases += 1; votes += int(want.sum()); flagged += st["flagged_cols"]; fixes += st["fixes"]; atomics += st["atomics"]
print("ok: %d cases, %.3g votes, %.3g atomics (%.2f per vote), flagged columns %d (%.2e per boundary), fixes %d" %
      (cases, votes, atomics, atomics / max(votes, 1), flagged, flagged / max(atomics, 1), fixes))
