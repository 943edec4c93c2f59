#!/usr/bin/env python
"""Small driver for ncu: runs the configs[3] flow (8-D PWQuad, 6 mask cells, 32 bins, [64]*3) forward in train-mode BN a few
times on 2^22 points.  Usage: python tools/profile_quad.py [log2_points]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nf_b200.normalizing_flows.manager import PWQuadManager  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 22)
torch.manual_seed(1234)
NF = PWQuadManager(n_flow=8)
NF.create_model(6, 32, [64] * 3)
model = NF._model.train()
x = torch.rand(n, 8, device="cuda", dtype=torch.float32)
with torch.no_grad():
    for _ in range(2):
        model(x)
torch.cuda.synchronize()
print("done quad")
