"""Loader for the host-side rasteriser harness (tests/hostsim.cpp). Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        out = os.path.join(HERE, "_hostsim")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libhostsim.so")
        srcs = [os.path.join(HERE, "hostsim.cpp"), os.path.join(ROOT, "rcvpose_b200", "csrc", "raster_core.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fvisibility=hidden", "-o", so, srcs[0]],
                           check=True)
        L = C.CDLL(so)
        L.hostsim_render.restype = C.c_int
        L.hostsim_render.argtypes = [C.c_void_p, C.c_void_p, C.c_long] + [C.c_int] * 6 + [C.c_void_p, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def render(p, R, D, tile=None, Dp=None, sqrt_perturb=0):
    """Render points into a tile (i0, ni, j0, nj); returns (volume view (ni,nj,D) int32, stats)."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.int32)
    i0, ni, j0, nj = tile if tile is not None else (0, D, 0, D)
    Dp = Dp or (D | 1)
    buf = np.zeros((ni, nj, Dp), dtype=np.int32)
    stats = np.zeros(8, dtype=np.int64)
    rc = lib().hostsim_render(p.ctypes.data, R.ctypes.data, p.shape[0], D, Dp, i0, ni, j0, nj, buf.ctypes.data, sqrt_perturb,
                              stats.ctypes.data)
    assert rc == 0, "rasteriser emitted a vote outside its tile"
    assert not buf[:, :, D:].any()
    return buf[:, :, :D], dict(votes=int(stats[0]), calls=int(stats[1]), ring_slices=int(stats[3]), dense_slices=int(stats[4]),
                               lane_tasks=int(stats[5]))
