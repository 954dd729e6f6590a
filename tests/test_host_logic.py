"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol that
include/rcvvote.h declares, refuses to run without an sm_100 GPU (no CPU fallback), and the
frame-sharding / result-gather logic of the multi-GPU path works over gloo with world_size 2."""
import ctypes
import os
import re
import socket

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "rcvvote.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rcv_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from rcvpose_b200 import _lib, build
    lib = build.build_library()
    L = ctypes.CDLL(lib)
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "include/rcvvote.h declares %s but librcvvote.so does not export it" % n
    assert sorted(_lib.EXPORTS) == names, "ctypes binding and header disagree"
    L.rcv_abi_version.restype = ctypes.c_int
    assert L.rcv_abi_version() == _lib.RCV_ABI_VERSION


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    from rcvpose_b200 import _lib, api
    with pytest.raises(Exception) as ei:
        api.VoteContext(0, max_items=4, max_points_total=1024, max_grid=64)
    assert "CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)
    from rcvpose_b200 import AccumulatorSpace as A
    with pytest.raises(Exception):
        A.Accumulator_3D(np.zeros((4, 3)), np.ones(4, np.float32))
    assert _lib.load().rcv_abi_version() == _lib.RCV_ABI_VERSION == 2


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under rcvpose_b200/ may import or load it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rcvpose_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "librcv_oracle" not in text, f


def test_shard_range_partitions_frames():
    from rcvpose_b200.pipeline import shard_range
    for n in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    from rcvpose_b200.pipeline import pack_results, unpack_results
    rng = np.random.default_rng(0)
    B, Kp = 5, 3
    c = torch.from_numpy(rng.normal(size=(B, Kp, 3)))
    RT = torch.from_numpy(rng.normal(size=(B, 4, 4)))
    peak = torch.from_numpy(rng.integers(0, 5000, size=(B, Kp)).astype(np.int32))
    st = torch.from_numpy(rng.integers(0, 32, size=(B, Kp)).astype(np.int32))
    rows = pack_results(c, RT, peak, st)
    assert rows.shape == (B, 3 * Kp + 16 + 2 * Kp)
    c2, RT2, p2, s2 = unpack_results(rows, Kp)
    assert torch.equal(c, c2) and torch.equal(RT, RT2) and torch.equal(peak, p2) and torch.equal(st, s2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from rcvpose_b200.pipeline import gather_results, pack_results, shard_range, unpack_results
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(n_frames, rank, world)
        f = torch.arange(lo, hi, dtype=torch.float64)
        centres = (f[:, None, None] * 10 + torch.arange(3, dtype=torch.float64)[None, :, None] + torch.arange(3, dtype=torch.float64)[None, None, :] / 10)
        RT = f[:, None, None] + torch.eye(4, dtype=torch.float64)[None]
        peak = (f[:, None] + torch.arange(3)[None, :]).to(torch.int32)
        st = torch.zeros((hi - lo, 3), dtype=torch.int32)
        rows = pack_results(centres, RT, peak, st)
        out_known = gather_results(rows, counts=[shard_range(n_frames, r, world)[1] - shard_range(n_frames, r, world)[0] for r in range(world)])
        out_probe = gather_results(rows)                      # counts discovered by a first all_gather
        c, R, p, s = unpack_results(out_known, 3)
        ok = bool(torch.equal(out_known, out_probe) and out_known.shape[0] == n_frames
                  and torch.equal(c[:, 0, 0], torch.arange(n_frames, dtype=torch.float64) * 10)
                  and torch.equal(p[:, 2], (torch.arange(n_frames) + 2).to(torch.int32))
                  and torch.equal(R[:, 3, 3], torch.arange(n_frames, dtype=torch.float64) + 1))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [8, 7])
def test_gather_results_world2_gloo(n_frames):
    """N>1 path on CPU: two ranks own contiguous (possibly unequal) frame ranges, results are gathered in frame order."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


_DROPIN_SNIPPET = r"""
import sys
sys.path.insert(0, {pkg!r})            # INTEGRATION.md section 1, verbatim: the package DIRECTORY goes first on sys.path
import AccumulatorSpace                  # reference train.py:12 / AccumulatorSpace.py imported by name
from AccumulatorSpace import estimate_6d_pose_lm, estimate_6d_pose_lmo, rgbd_to_point_cloud, Accumulator_3D, linemod_K, read_depth
from util.horn import HornPoseFitting    # reference AccumulatorSpace.py:2
assert AccumulatorSpace.__file__.startswith({pkg!r}), AccumulatorSpace.__file__
assert callable(HornPoseFitting().lmshorn) and linemod_K.shape == (3, 3)
{body}
print("DROPIN_OK")
"""


def _run_dropin(body=""):
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "rcvpose_b200")
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    res = subprocess.run([sys.executable, "-c", _DROPIN_SNIPPET.format(pkg=pkg, body=body)], cwd="/tmp", env=env, capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0 and "DROPIN_OK" in res.stdout, res.stdout + res.stderr


def test_sys_path_dropin_imports_like_the_reference():
    """with rcvpose_b200/ first on sys.path the reference's own import lines resolve to the shims
    (no parent package: the shims must not depend on relative imports)."""
    _run_dropin()


@pytest.mark.gpu
def test_sys_path_dropin_runs_the_reference_shaped_loop():
    """The reference's per-image body (AccumulatorSpace.py:606-662: cloud, /1000, Accumulator_3D per keypoint, lmshorn) through the
    top-level shims, against the oracle."""
    body = r'''
import numpy as np
sys.path.append({root!r})
from oracle import oracle
from rcvpose_b200 import synth
fr = synth.config3_frame(3)
K = fr["K"]; est = np.zeros((3, 3)); want = np.zeros((3, 3))
for k in range(3):
    radial_est = np.where(fr["radius"][k] <= fr["max_radii_dm"][k], fr["radius"][k], 0)     # :612-616, npy branch
    depth_map = fr["depth"] * (radial_est != 0)
    xyz_mm = rgbd_to_point_cloud(K, depth_map)
    assert np.array_equal(xyz_mm, oracle.rgbd_to_point_cloud(K, depth_map))
    radial_list = radial_est[depth_map.nonzero()]
    xyz = xyz_mm / 1000
    est[k] = Accumulator_3D(xyz, radial_list)[0]
    want[k] = oracle.Accumulator_3D(xyz, radial_list)[0]
assert np.array_equal(est, want), (est, want)
P1 = fr["kpts_mm"]
RT = np.zeros((4, 4)); RTo = np.zeros((4, 4))
HornPoseFitting().lmshorn(P1, est, 3, RT)
oracle.lmshorn(P1, want, 3, RTo)
assert np.allclose(RT, RTo, atol=1e-9), (RT, RTo)
'''.format(root=ROOT)
    _run_dropin(body)


def test_shard_by_cost_balances_contiguous_ranges():
    """YCB-shaped work is balanced by the sum N*R^2 cost model, not by frame count."""
    from rcvpose_b200 import pipeline, synth
    pr = synth.frame_params(2048, 3, seed=3, obj_radius_mm=(60.0, 110.0), approach=True)
    cost = synth.frame_cost(pr, synth.ycb_K)
    assert cost.max() / cost.min() > 10                       # "per-item cost varies > 10x"
    for world in (1, 2, 4, 8):
        rg = pipeline.shard_by_cost(cost, world)
        assert rg[0][0] == 0 and rg[-1][1] == len(cost) and all(rg[r][1] == rg[r + 1][0] for r in range(world - 1))
        sums = np.array([cost[a:b].sum() for a, b in rg])
        eq = np.array([cost[slice(*pipeline.shard_range(len(cost), r, world))].sum() for r in range(world)])
        assert sums.max() / sums.mean() < 1.01
        if world > 1:
            assert eq.max() / eq.mean() > 1.15                # the equal-count split of an approach sequence is not balanced
    assert pipeline.shard_by_cost([], 3) == [(0, 0)] * 3
    assert pipeline.shard_by_cost([0.0, 0.0, 0.0, 0.0], 2) == [(0, 2), (2, 4)]
    rg = pipeline.shard_by_cost([5.0], 4)
    assert sum(b - a for a, b in rg) == 1


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """include/rcvvote.h is the boundary a reference-side binding compiles against: it must be valid C99 on its own (no C++-isms,
    no torch types) and a C program must link against librcvvote.so; without a GPU rcv_create reports RCV_E_NOGPU."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from rcvpose_b200 import _lib
    _lib.load()
    src = tmp_path / "c_abi.c"
    src.write_text('''
#include "rcvvote.h"
#include <stdio.h>
int main(void) {
  rcv_config cfg = {RCV_ABI_VERSION, 16, 1 << 20, 256, 0, 0, 0, 0};
  rcv_ctx* ctx = 0;
  int rc = rcv_create(0, &cfg, &ctx);
  printf("%d %d %d\\n", rcv_abi_version(), rc, ctx != 0);
  if (ctx) rcv_destroy(ctx);
  return 0;
}
''')
    exe = tmp_path / "c_abi"
    libdir = os.path.join(ROOT, "rcvpose_b200")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src),
                          "-L", libdir, "-lrcvvote", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    abi, rc, has_ctx = [int(x) for x in out.stdout.split()]
    assert abi == _lib.RCV_ABI_VERSION
    import torch
    if torch.cuda.is_available():
        assert rc == 0 and has_ctx == 1
    else:
        assert rc == -4 and has_ctx == 0          # RCV_E_NOGPU: there is no CPU fallback
