#!/usr/bin/env python
"""Times one variance-loss training step (forward + fused backward) of the cfg2 flow.  Development aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nf_b200.normalizing_flows.manager import PWLinManager, PWQuadManager  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 16)
torch.manual_seed(1234)
kind = sys.argv[2] if len(sys.argv) > 2 else "lin"
if kind == "lin":
    NF = PWLinManager(n_flow=8)
    NF.create_model(4, 6, 32, [64] * 3, 4)
else:                                   # cfg4's flow: 8-D PWQuad, 6 mask cells, 32 bins
    NF = PWQuadManager(n_flow=8)
    NF.create_model(6, 32, [64] * 3)
model = NF._model.train()
x = torch.rand(n, 8, device="cuda", dtype=torch.float32)
f = torch.exp(-((x - 0.5) ** 2).sum(-1) / 0.2)


def step():
    model.zero_grad()
    XJ = model(x)
    loss = torch.var(f * XJ[:, -1])
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
for name, fn in (("fwd+bwd", step), ("fwd only", lambda: model(x))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(kind + " %s  B=%d: %.3f ms  %.3e points/s" % (name, n, ms, n / ms * 1e3))
