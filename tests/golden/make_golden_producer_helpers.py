"""Seeded RadiusTrunk + input shared by tests/golden/make_golden_producer.py and tests/test_producer.py (no reference import)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rcvpose_b200 import producer  # noqa: E402


def seeded_trunk(seed=1234):
    torch.manual_seed(seed)
    net = producer.RadiusTrunk().eval()
    g = torch.Generator().manual_seed(seed + 1)
    for k, v in net.state_dict().items():
        if k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        if k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
    x = torch.randn((2, 3, 64, 96), generator=g)
    return net, x
