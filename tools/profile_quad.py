#!/usr/bin/env python
"""ncu driver: cfg4's flow (8-D PWQuad, 6 mask cells, 32 bins, [64]*3) forward on 2^n points, eval or train BN."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nf_b200.normalizing_flows.manager import PWQuadManager  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "eval"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 21)
torch.manual_seed(1234)
NF = PWQuadManager(n_flow=8)
NF.create_model(6, 32, [64] * 3)
model = NF._model.train(mode == "train")
x = torch.rand(n, 8, device="cuda", dtype=torch.float32)
with torch.no_grad():
    for _ in range(2):
        model(x)
torch.cuda.synchronize()
print("done", mode)
