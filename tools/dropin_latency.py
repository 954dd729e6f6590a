"""Zero-change drop-in, timed: the reference's per-image loop (AccumulatorSpace.py:594-662 in shape: per keypoint mask the depth,
rgbd_to_point_cloud, Accumulator_3D; then HornPoseFitting.lmshorn) calling the shims by the reference's module names with this
package's directory first on sys.path.  Host arrays in and out, one synchronising call after the other -- what a user of the
reference gets without touching their code.  Prints one JSON line."""
import json
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rcvpose_b200"))
import numpy as np
import AccumulatorSpace as A                       # the shim, imported the way the reference imports its own module
from util.horn import HornPoseFitting
sys.path.append(ROOT)
from rcvpose_b200 import synth

N = int(os.environ.get("RCV_DROPIN_FRAMES", "24"))
frames = [synth.config3_frame(200 + f) for f in range(N)]
horn = HornPoseFitting()


def one(fr):
    est = np.zeros((3, 3))
    for k in range(3):
        r = fr["radius"][k]
        dm = fr["depth"] * np.where(r != 0, 1, 0).astype(fr["depth"].dtype)
        xyz_mm = A.rgbd_to_point_cloud(A.linemod_K, dm)
        est[k] = A.Accumulator_3D(xyz_mm / 1000, r[dm.nonzero()])[0]
    RT = np.zeros((4, 4))
    horn.lmshorn(fr["kpts_mm"] - fr["centre_mm"], est, 3, RT)
    return est, RT


one(frames[0])
t0 = time.perf_counter()
errs = []
for fr in frames:
    est, RT = one(fr)
    errs.append(float(np.abs(est - fr["kpts_mm"]).max()))
dt = time.perf_counter() - t0
print(json.dumps({"tool": "dropin_latency", "frames": N, "keypoints": 3, "ms_per_frame": round(dt / N * 1e3, 3), "frames_per_s": round(N / dt, 1),
                  "max_keypoint_error_mm": round(max(errs), 3), "path": "sys.path drop-in: AccumulatorSpace.rgbd_to_point_cloud + Accumulator_3D x3 + util.horn.HornPoseFitting.lmshorn, host arrays, one image at a time"}))
