"""Drop-in for the reference's util/horn.py (HornPoseFitting.lmshorn, :75-181) on the GPU."""
import numpy as np

try:
    from ..AccumulatorSpace import _ctx
except ImportError:
    # Zero-change drop-in: rcvpose_b200/ itself is on sys.path, so this is the top-level package `util`
    # (reference AccumulatorSpace.py:2, `from util.horn import HornPoseFitting`).
    import os as _os
    import sys as _sys
    _root = _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
    if _root not in _sys.path:
        _sys.path.append(_root)
    from rcvpose_b200.AccumulatorSpace import _ctx


class HornPoseFitting:
    def __init__(self):
        super(HornPoseFitting, self).__init__()

    def lmshorn(self, P1, P2, n, A):
        """Fills the caller's 4x4 `A` in place with [R|T] minimising sum |R*P1_i + T - P2_i|^2 and
        returns None; P1/P2 are left unchanged (the reference centres and then restores them)."""
        ctx = _ctx()
        P1 = np.asarray(P1, dtype=np.float64)[:n]
        P2 = np.asarray(P2, dtype=np.float64)[:n]
        RT = ctx.horn_batch_host(P1.reshape(n, 3), P2.reshape(1, n, 3))
        A[...] = RT[0]
        return None
