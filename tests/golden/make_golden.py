"""Generate golden fixtures by running the REAL reference (aaronWool/rcvpose) in this container.

Usage (build container only -- /root/reference does not exist on the GPU box):
    NUMBA_NUM_THREADS=1 python tests/golden/make_golden.py

Recipe = SURVEY.md Appendix A: stub the three absent, unused modules, import AccumulatorSpace and
util.horn unmodified, force single-threaded numba (the reference's prange vote loop is racy --
SURVEY section 0.2), run on seeded synthetic inputs and store inputs' seeds + outputs.
Nothing from the reference's source is copied; only its *outputs* are stored.
"""
import os
import sys
import types

os.environ.setdefault("NUMBA_NUM_THREADS", "1")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

for name in ("open3d", "h5py", "matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")
import numba  # noqa: E402
import AccumulatorSpace as A  # noqa: E402
from util.horn import HornPoseFitting  # noqa: E402

from rcvpose_b200 import synth  # noqa: E402

assert numba.get_num_threads() == 1, "run with NUMBA_NUM_THREADS=1"


def ref_volume(xyz, radial_list):
    """Re-run the reference's prelude lines on copies and call the reference's own fast_for to get
    the full volume (Accumulator_3D only returns the centre)."""
    acc_unit = 5
    xyz_mm = xyz * 1000 / acc_unit
    m = [np.mean(xyz_mm[:, c]) for c in range(3)]
    for c in range(3):
        xyz_mm[:, c] -= m[c]
    r_mm = radial_list * 100 / acc_unit
    zb = int(xyz_mm.min() - r_mm.max()) + 1
    if zb < 0:
        xyz_mm -= zb
    length = int(xyz_mm.max())
    D = length + int(r_mm.max())
    V = A.fast_for(xyz_mm, r_mm, np.zeros((D, D, D)))
    return V, zb, D, np.array(m), xyz_mm


def case_from_frame(fr, k):
    xyz_mm = A.rgbd_to_point_cloud(fr["K"], fr["depth"] * (fr["radius"][k] != 0))
    xyz, rl = synth.frame_to_points(fr["K"], fr["depth"], fr["radius"][k])
    assert np.array_equal(xyz_mm / 1000, xyz)
    centre = A.Accumulator_3D(xyz.copy(), rl.copy())
    V, zb, D, mean, p = ref_volume(xyz.copy(), rl.copy())
    Vi = V.astype(np.int32)
    assert np.array_equal(Vi, V)
    am = np.argwhere(V == V.max())
    return dict(n=xyz.shape[0], xyz_mm_first=xyz_mm[:4].copy(), xyz_mm_xor=np.bitwise_xor.reduce(np.ascontiguousarray(xyz_mm).view(np.uint64), axis=0), centre_mm=centre[0].copy(),
                ties=centre.shape[0], zb=zb, D=D, mean=mean, votes=int(V.sum()), peak=int(V.max()), argmax=am[0].copy(),
                vol_crc=np.array([int(np.bitwise_xor.reduce((Vi.ravel().astype(np.int64) * (np.arange(Vi.size, dtype=np.int64) % 1000003 + 1)) & 0xFFFFFFFF))]),
                vol_nonzero=int((Vi != 0).sum()), volume=Vi, p_first=p[:4].copy())


def main():
    out = {}
    # --- config 1: the survey's headline frame (full volume stored, compressed) ---
    fr = synth.config1_frame()
    c = case_from_frame(fr, 0)
    c["xyz_mm"] = A.rgbd_to_point_cloud(fr["K"], fr["depth"] * (fr["radius"][0] != 0))
    for k, v in c.items():
        out["c1_" + k] = v
    print("config1: n=%d D=%d zb=%d votes=%d peak=%d centre=%s" % (c["n"], c["D"], c["zb"], c["votes"], c["peak"], c["centre_mm"]))
    # --- config 3 frames 0..3, 3 keypoints each: summaries + a sparse sample of the volume ---
    for f in range(4):
        fr = synth.config3_frame(f)
        for k in range(3):
            c = case_from_frame(fr, k)
            V = c.pop("volume")
            rng = np.random.default_rng(1000 + 10 * f + k)
            sel = rng.integers(0, V.size, size=4096)
            c["sample_idx"] = sel
            c["sample_val"] = V.ravel()[sel]
            c["slice_sums"] = V.sum(axis=(1, 2)).astype(np.int64)
            for kk, v in c.items():
                out["c3_f%d_k%d_%s" % (f, k, kk)] = v
            print("config3 f=%d k=%d: n=%d D=%d votes=%d peak=%d" % (f, k, c["n"], c["D"], c["votes"], c["peak"]))
    # --- small random point clouds through Accumulator_3D (float32 and float64 radii, zb>=0 and <0) ---
    rng = np.random.default_rng(7)
    for t in range(8):
        n = int(rng.integers(1, 60))
        xyz = rng.normal(0, 0.02 + 0.01 * t, size=(n, 3)) + rng.normal(0, 0.3, size=(1, 3))
        kp = xyz.mean(0) + rng.normal(0, 0.05, size=3)
        rl = (np.linalg.norm(xyz - kp, axis=1) * 10 + rng.normal(0, 0.02, n))
        rl = rl.astype(np.float32) if t % 2 == 0 else rl.astype(np.float64)
        centre = A.Accumulator_3D(xyz.copy(), rl.copy())
        V, zb, D, mean, p = ref_volume(xyz.copy(), rl.copy())
        out["r%d_xyz" % t] = xyz
        out["r%d_rl" % t] = rl
        out["r%d_centre" % t] = centre
        out["r%d_zb" % t] = zb
        out["r%d_D" % t] = D
        out["r%d_mean" % t] = mean
        out["r%d_volume" % t] = V.astype(np.int32)
        out["r%d_p" % t] = p
    # --- Horn: exact rigid motions and noisy triples ---
    horn = HornPoseFitting()
    for t in range(6):
        P1 = rng.normal(0, 60, size=(3 if t < 4 else 5, 3))
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        w, x, y, z = q
        Rm = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                       [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                       [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        T = rng.normal(0, 300, size=3)
        P2 = P1 @ Rm.T + T + (rng.normal(0, 2.5, size=P1.shape) if t % 2 else 0.0)
        RT = np.zeros((4, 4))
        a, b = P1.copy(), P2.copy()
        horn.lmshorn(a, b, P1.shape[0], RT)
        out["h%d_P1" % t] = P1
        out["h%d_P2" % t] = P2
        out["h%d_RT" % t] = RT
    # --- numpy pairwise mean pins (strided column view, like xyz_mm[:,c]) ---
    for t, n in enumerate([1, 7, 8, 9, 127, 128, 129, 1000, 3189, 40570]):
        a = rng.normal(0, 50, size=(n, 3)) + 17.0
        out["pw%d_in" % t] = a
        out["pw%d_mean" % t] = np.array([np.mean(a[:, c]) for c in range(3)])
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_golden.npz"))


if __name__ == "__main__":
    main()
