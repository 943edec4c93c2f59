#!/usr/bin/env python
"""Training-step time of the cfg2 flow at small minibatches (where the kernel choice flips).  Development aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nf_b200.normalizing_flows.manager import PWLinManager  # noqa: E402

torch.manual_seed(1234)
NF = PWLinManager(n_flow=8)
NF.create_model(4, 6, 32, [64] * 3, 4)
model = NF._model.train()
for n in (256, 512, 1000, 2000, 2048, 4096):
    x = torch.rand(n, 8, device="cuda", dtype=torch.float32)
    f = torch.exp(-((x - 0.5) ** 2).sum(-1) / 0.2)

    def step():
        model.zero_grad()
        XJ = model(x)
        torch.var(f * XJ[:, -1]).backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    print("B=%5d fwd+bwd %.3f ms" % (n, e0.elapsed_time(e1) / 10))
