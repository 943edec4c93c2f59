#!/usr/bin/env python
"""In-situ kernel durations (torch.profiler / CUPTI, no replay) of one variance-loss training step.
Usage: python tools/insitu_step.py {cfg5|cfg2} [log2_points]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from nf_b200.normalizing_flows.manager import PWLinManager, PWQuadManager  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 16)
torch.manual_seed(1234)
if which == "cfg5":
    NF = PWQuadManager(n_flow=16)
    NF.create_model(8, 64, [256] * 4)
    d = 16
else:
    NF = PWLinManager(n_flow=8)
    NF.create_model(4, 6, 32, [64] * 3, 4)
    d = 8
model = NF._model.train()
x = torch.rand(n, d, device="cuda", dtype=torch.float32)
f = torch.exp(-((x - 0.5) ** 2).sum(-1) / 0.4)


def step():
    model.zero_grad()
    XJ = model(x)
    torch.var(f * XJ[:, -1]).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_time > 0:
        a = agg.setdefault(e.name[:90], [0, 0.0])
        a[0] += 1
        a[1] += e.device_time
tot = sum(a[1] for a in agg.values())
print("%s, 2^%d points: total kernel time per step %.3f ms" % (which, n.bit_length() - 1, tot / 3 / 1e3))
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print("%-92s n/step=%6.1f  avg=%8.1f us  share=%5.1f%%" % (name, a[0] / 3, a[1] / a[0], 100 * a[1] / tot))
