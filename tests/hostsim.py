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
        L.hostsim_render_v6.restype = C.c_int
        L.hostsim_render_v6.argtypes = [C.c_void_p, C.c_void_p, C.c_long] + [C.c_int] * 6 + [C.c_void_p, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


def render(p, R, D, tile=None, Dp=None, sqrt_perturb=0, slab=32):
    """Render points into the REFERENCE-axis tile (i0, ni, j0, nj); returns (volume view (ni,nj,D) int32, stats).
    The kernel's internal axes are (y,x,z): the reference tile is the internal slab A in [j0,j0+nj) with rows B in
    [i0,i0+ni).  The kernel never builds slabs of more than 32 slices (one mask bit per slice in the polar pass),
    so the request is rendered in slabs of `slab` <= 32 slices and stitched back together."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.int32)
    i0, ni, j0, nj = tile if tile is not None else (0, D, 0, D)
    Dp = Dp or (D | 1)
    stats = np.zeros(10, dtype=np.int64)
    parts, tot = [], np.zeros(10, dtype=np.int64)
    for a0 in range(j0, j0 + nj, slab):
        na = min(slab, j0 + nj - a0)
        part = np.zeros((na, ni, Dp), dtype=np.int32)
        rc = lib().hostsim_render_v6(p.ctypes.data, R.ctypes.data, p.shape[0], D, Dp, a0, na, i0, ni, part.ctypes.data, sqrt_perturb,
                                     stats.ctypes.data)
        assert rc == 0, "rasteriser emitted a vote outside its tile (rc=%d)" % rc
        parts.append(part)
        tot += stats
    stats = tot
    buf = np.concatenate(parts, axis=0)
    assert not buf[:, :, D:].any()
    buf = np.ascontiguousarray(buf.transpose(1, 0, 2))
    return buf[:, :, :D], dict(votes=int(stats[0]), calls=int(stats[1]), ring_chunks=int(stats[3]), dense_slices=int(stats[4]),
                               lane_tasks=int(stats[5]), polar_cells=int(stats[7]), slow_calls=int(stats[6]), ring2_tasks=int(stats[8]))
