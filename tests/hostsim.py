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


_LIB_RUNS = None


def lib_runs():
    """Harness of the run-length rasteriser (tests/hostsim_runs.cpp over rcvpose_b200/csrc/runs_core.h)."""
    global _LIB_RUNS
    if _LIB_RUNS is None:
        out = os.path.join(HERE, "_hostsim")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libhostsim_runs.so")
        srcs = [os.path.join(HERE, "hostsim_runs.cpp"), os.path.join(ROOT, "rcvpose_b200", "csrc", "runs_core.h"),
                os.path.join(ROOT, "rcvpose_b200", "csrc", "raster_core.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fvisibility=hidden", "-o", so, srcs[0]], check=True)
        L = C.CDLL(so)
        L.hostsim_runs_render.restype = C.c_int
        L.hostsim_runs_render.argtypes = [C.c_void_p, C.c_void_p, C.c_long] + [C.c_int] * 9 + [C.c_void_p, C.c_int, C.c_void_p]
        _LIB_RUNS = L
    return _LIB_RUNS


def render_runs(p, R, D, glo=4, ghi=4, slab=4, NC=None, band=None, clip=False, sqrt_perturb=0):
    """Vote counts (D,D,D) int32 in REFERENCE axis order from the run-length rasteriser, rendered slab by slab like the kernel
    (slabs of `slab` y-slices; band = (j0, nj) = the x rows of every tile: the grid's rows by default, as in the kernel -- a
    lane only draws the columns of its range that fall on the tile's rows; (-glo, D + glo + ghi) adds guard rows)."""
    p = np.ascontiguousarray(p, dtype=np.float64)
    R = np.ascontiguousarray(R, dtype=np.int32)
    j0, nj = band if band is not None else (0, D)
    NC = NC or min(slab, 4)
    tot = np.zeros(8, dtype=np.int64)
    vol = np.zeros((D, D, D), dtype=np.int32)     # [x][y][z]
    for a0 in range(0, D, slab):
        na = min(slab, D - a0)
        part = np.zeros((na, nj, D), dtype=np.int32)
        stats = np.zeros(8, dtype=np.int64)
        rc = lib_runs().hostsim_runs_render(p.ctypes.data, R.ctypes.data, p.shape[0], D, glo, ghi, a0, na, j0, nj, NC, int(clip),
                                            part.ctypes.data, sqrt_perturb, stats.ctypes.data)
        assert rc == 0, "run rasteriser addressed a cell outside its tile (rc=%d, oob=%d)" % (rc, stats[5])
        tot += stats
        jlo, jhi = max(j0, 0), min(j0 + nj, D)
        vol[jlo:jhi, a0:a0 + na, :] = part[:, jlo - j0:jhi - j0, :].transpose(1, 0, 2)
    return vol, dict(cols=int(tot[0]), live_cols=int(tot[1]), flagged_cols=int(tot[2]), exact_calls=int(tot[3]), fixes=int(tot[4]),
                     atomics=int(tot[6]))


_LIB_HORN = None


def horn(P1, P2):
    """rcvpose_b200/csrc/horn_core.h compiled for the host (tests/hostsim_horn.cpp): 4x4 [R|T] with R P1_i + T ~ P2_i."""
    global _LIB_HORN
    if _LIB_HORN is None:
        out = os.path.join(HERE, "_hostsim")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "libhostsim_horn.so")
        srcs = [os.path.join(HERE, "hostsim_horn.cpp"), os.path.join(ROOT, "rcvpose_b200", "csrc", "horn_core.h")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fvisibility=hidden", "-o", so, srcs[0]], check=True)
        _LIB_HORN = C.CDLL(so)
        _LIB_HORN.hostsim_horn.restype = None
        _LIB_HORN.hostsim_horn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    P1 = np.ascontiguousarray(P1, dtype=np.float64); P2 = np.ascontiguousarray(P2, dtype=np.float64)
    RT = np.zeros((4, 4))
    _LIB_HORN.hostsim_horn(P1.ctypes.data, P2.ctypes.data, P1.shape[0], RT.ctypes.data)
    return RT
