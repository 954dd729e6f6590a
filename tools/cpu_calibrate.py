"""Times the REAL reference fast_for (numba, imported from /root/reference -- build container only) next to the oracle's
ports on the same cores, on the config-1 frame.  Calibrates bench.py's `cpu_baseline` / `--impl reference` (kind "port")."""
import json, os, sys, time, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
for name in ("open3d", "h5py", "matplotlib", "matplotlib.pyplot"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")
import warnings; warnings.filterwarnings("ignore")
import AccumulatorSpace as A
from oracle import oracle
from rcvpose_b200 import synth
fr = synth.config1_frame()
xyz, rl = synth.frame_to_points(fr["K"], fr["depth"], fr["radius"][0])
pre = oracle.prelude(xyz, rl)
p, R, D = pre["p"], pre["R"], pre["D"]
rv = rl * 100 / 5
cores = len(os.sched_getaffinity(0))
def t(f, reps=3):
    f(); b = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); f(); b = min(b, time.perf_counter() - t0)
    return b
t_ref = t(lambda: A.fast_for(p, rv, np.zeros((D, D, D))))
t_full = t(lambda: oracle.fast_for(p, R, D, method="full"))
t_cull = t(lambda: oracle.fast_for(p, R, D, method="brute"))
print(json.dumps({"cores": cores, "N": int(p.shape[0]), "D": int(D), "tests": int(p.shape[0]) * D ** 3,
                  "reference_numba_fast_for_s": t_ref, "port_full_s": t_full, "port_culled_s": t_cull,
                  "reference_tests_per_s": p.shape[0] * D ** 3 / t_ref, "port_full_tests_per_s": p.shape[0] * D ** 3 / t_full}))
