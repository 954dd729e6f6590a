import sys, numpy as np
sys.path.insert(0, '.')
from rcvpose_b200 import AccumulatorSpace as A, synth
from oracle import oracle
fr = synth.config1_frame()
xyz, rl = synth.frame_to_points(fr["K"], fr["depth"], fr["radius"][0])
for sel in ([2582], [137], [2291]):
    # keep the grid of the full frame: all points, but zero radius (no votes) for the others
    r2 = np.zeros_like(rl); r2[sel] = rl[sel]
    # points with radius 0 vote nothing but keep the prelude (mean, D) identical... rmax changes D; use oracle on same input
    vol, out = A.vote_volume(xyz, r2)
    want, info = oracle.Accumulator_3D(xyz, r2, return_info=True, return_volume=True)
    d = vol.astype(np.int64) - info["volume"].astype(np.int64)
    bad = np.argwhere(d != 0)
    print(sel, "D", info["D"], "votes", int(info["volume"].sum()), "got", int(vol.sum()), "mismatches", bad.tolist(), d[d != 0].tolist())
