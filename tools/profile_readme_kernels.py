#!/usr/bin/env python
"""Kernel table of the README example (BASELINE configs[0]) as it runs now (graph epochs): torch.profiler over the last
epochs of a 60-epoch run.  Development aid."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402
from nf_b200.normalizing_flows.manager import PWQuadManager  # noqa: E402


def camel(x):
    return torch.exp(-((x[:, 0] - 0.75) ** 2 + (x[:, 1] - 0.75) ** 2) / (0.2 ** 2)) + \
        torch.exp(-((x[:, 0] - 0.25) ** 2 + (x[:, 1] - 0.25) ** 2) / (0.2 ** 2))


def run(epochs):
    torch.manual_seed(0)
    NF = PWQuadManager(n_flow=2)
    NF.create_model(2, 4, [3] * 3)
    optim = torch.optim.Adamax(NF._model.parameters(), lr=2e-3, weight_decay=1e-04)
    NF._train_variance_forward_seq(camel, optim, True, tempfile.mkdtemp(), 10000, epochs, 0, False, True, preburn_time=50)
    torch.cuda.synchronize()


run(20)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    run(40)
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = {}
for e in ev:
    t = tot.setdefault(e.name[:70], [0, 0.0])
    t[0] += 1
    t[1] += e.device_time
all_us = sum(v[1] for v in tot.values())
print("40 epochs: %d kernels, %.1f ms of device time = %.0f us per epoch" % (len(ev), all_us / 1e3, all_us / 40))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:18]:
    print("%-72s n=%5d  avg=%7.1f us  total=%7.1f us/epoch" % (k, v[0], v[1] / v[0], v[1] / 40))
