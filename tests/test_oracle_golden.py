"""Pins the CPU oracle (oracle/rcv_oracle.c) against the real reference's outputs.

The fixtures in tests/golden/reference_golden.npz were produced by importing the unmodified
reference (AccumulatorSpace.py, util/horn.py) single-threaded -- see tests/golden/make_golden.py.
Integer results must match bit-for-bit; float64 results to the last ulp (rtol 0 where the
operation order is reproduced exactly, 1e-12 for Horn where Python scalars are involved).
"""
import numpy as np
import pytest

from oracle import oracle
from rcvpose_b200 import synth


def test_pairwise_mean_matches_numpy(golden):
    for t in range(10):
        a = golden["pw%d_in" % t]
        for c in range(3):
            got = oracle.lib().orc_pairwise_sum(a[:, c].ctypes.data, a.shape[0], 3) / a.shape[0]
            assert got == golden["pw%d_mean" % t][c]
            assert got == np.mean(a[:, c])


def test_config1_full_volume_bit_exact(golden):
    fr = synth.config1_frame()
    xyz, rl = synth.frame_to_points(fr["K"], fr["depth"], fr["radius"][0])
    assert xyz.shape[0] == int(golden["c1_n"]) == 3189
    xyz_mm = oracle.rgbd_to_point_cloud(fr["K"], fr["depth"] * (fr["radius"][0] != 0))
    assert np.array_equal(xyz_mm[:4], golden["c1_xyz_mm_first"])
    assert np.array_equal(xyz_mm, golden["c1_xyz_mm"])
    assert np.array_equal(np.bitwise_xor.reduce(xyz_mm.view(np.uint64), axis=0), golden["c1_xyz_mm_xor"])
    pre = oracle.prelude(xyz, rl)
    assert pre["D"] == int(golden["c1_D"]) == 86
    assert pre["zb"] == int(golden["c1_zb"]) == -43
    assert np.array_equal(pre["mean"], golden["c1_mean"])
    assert np.array_equal(pre["p"][:4], golden["c1_p_first"])
    vol = oracle.fast_for(pre["p"], pre["R"], pre["D"])
    assert np.array_equal(vol, golden["c1_volume"])          # every voxel of the reference's volume
    assert int(vol.sum()) == int(golden["c1_votes"])
    idx, mx, ties = oracle.peak(vol)
    assert np.array_equal(idx, golden["c1_argmax"]) and mx == int(golden["c1_peak"]) and ties == int(golden["c1_ties"])
    c = oracle.center_mm(idx, pre["zb"], pre["mean"])
    assert np.array_equal(c, golden["c1_centre_mm"])
    # scatter renderer == brute force
    assert np.array_equal(oracle.fast_for(pre["p"], pre["R"], pre["D"], method="scatter"), vol)


@pytest.mark.parametrize("f", range(4))
def test_config3_frames(golden, f):
    fr = synth.config3_frame(f)
    for k in range(3):
        g = lambda name: golden["c3_f%d_k%d_%s" % (f, k, name)]
        xyz, rl = synth.frame_to_points(fr["K"], fr["depth"], fr["radius"][k])
        assert xyz.shape[0] == int(g("n"))
        xyz_mm = oracle.rgbd_to_point_cloud(fr["K"], fr["depth"] * (fr["radius"][k] != 0))
        assert np.array_equal(np.bitwise_xor.reduce(xyz_mm.view(np.uint64), axis=0), g("xyz_mm_xor"))
        pre = oracle.prelude(xyz, rl)
        assert pre["D"] == int(g("D")) and pre["zb"] == int(g("zb"))
        assert np.array_equal(pre["mean"], g("mean"))
        vol = oracle.fast_for(pre["p"], pre["R"], pre["D"], method="scatter")
        assert int(vol.sum()) == int(g("votes"))
        assert np.array_equal(vol.ravel()[g("sample_idx")], g("sample_val"))
        assert np.array_equal(vol.sum(axis=(1, 2)), g("slice_sums"))
        assert int((vol != 0).sum()) == int(g("vol_nonzero"))
        idx, mx, ties = oracle.peak(vol)
        assert np.array_equal(idx, g("argmax")) and mx == int(g("peak"))
        assert np.array_equal(oracle.center_mm(idx, pre["zb"], pre["mean"]), g("centre_mm"))


@pytest.mark.parametrize("t", range(8))
def test_random_clouds_full_path(golden, t):
    xyz, rl = golden["r%d_xyz" % t], golden["r%d_rl" % t]
    pre = oracle.prelude(xyz, rl)
    assert pre["D"] == int(golden["r%d_D" % t]) and pre["zb"] == int(golden["r%d_zb" % t])
    assert np.array_equal(pre["p"], golden["r%d_p" % t])
    vol = oracle.fast_for(pre["p"], pre["R"], pre["D"])
    assert np.array_equal(vol, golden["r%d_volume" % t])
    assert np.array_equal(oracle.fast_for(pre["p"], pre["R"], pre["D"], method="literal"), vol)
    assert np.array_equal(oracle.fast_for(pre["p"], pre["R"], pre["D"], method="scatter"), vol)
    c = oracle.Accumulator_3D(xyz, rl)
    assert np.array_equal(c[0], golden["r%d_centre" % t][0])


@pytest.mark.parametrize("t", range(6))
def test_horn_matches_reference(golden, t):
    P1, P2, RT = golden["h%d_P1" % t], golden["h%d_P2" % t], golden["h%d_RT" % t]
    A = np.zeros((4, 4))
    a, b = P1.copy(), P2.copy()
    oracle.lmshorn(a, b, P1.shape[0], A)
    assert np.array_equal(a, P1) and np.array_equal(b, P2)
    np.testing.assert_allclose(A, RT, rtol=0, atol=1e-12 * max(1.0, np.abs(RT).max()))
    assert abs(np.linalg.det(A[:3, :3]) - 1.0) < 1e-12


def test_empty_input_raises():
    with pytest.raises(ValueError):
        oracle.Accumulator_3D(np.zeros((0, 3)), np.zeros((0,), np.float32))
