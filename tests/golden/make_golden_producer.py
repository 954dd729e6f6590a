"""Golden fixture for the PyTorch input stage (rcvpose_b200/producer.py): outputs of the REAL reference's DenseFCNResNet152
(models/fcnresnet.py, imported unmodified from /root/reference) for a seeded state_dict and input.

Usage (build container only):  python tests/golden/make_golden_producer.py
The state_dict is the one a seeded `RadiusTrunk()` is born with (plus seeded BatchNorm statistics), loaded into the reference
model with load_state_dict(strict=True) -- which also proves that the parameter names and shapes coincide."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from models.fcnresnet import DenseFCNResNet152  # noqa: E402
from rcvpose_b200 import producer  # noqa: E402


from make_golden_producer_helpers import seeded_trunk  # noqa: E402


if __name__ == "__main__":
    net, x = seeded_trunk()
    ref = DenseFCNResNet152(3, 2).eval()
    ref.load_state_dict(net.state_dict(), strict=True)
    with torch.no_grad():
        seg, rad = ref(x.clone())
    keys = "\n".join("%s %s" % (k, tuple(v.shape)) for k, v in ref.state_dict().items())
    np.savez_compressed(os.path.join(HERE, "producer_golden.npz"), seg=seg.numpy(), radial=rad.numpy(), keys=np.array(keys),
                        x_checksum=np.array(float(x.double().sum())))
    print("wrote producer_golden.npz", seg.shape, float(seg.abs().max()), float(rad.abs().max()), len(keys.splitlines()), "tensors")
