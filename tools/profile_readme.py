#!/usr/bin/env python
"""Where the README example's (BASELINE configs[0]) wall time goes: cProfile of 60 epochs (host side) and a
torch.profiler kernel table of 10 epochs (device side).  Development aid."""
import cProfile
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nf_b200.normalizing_flows.manager import PWQuadManager  # noqa: E402


def camel(x):
    return torch.exp(-((x[:, 0] - 0.75) ** 2 + (x[:, 1] - 0.75) ** 2) / (0.2 ** 2)) + \
        torch.exp(-((x[:, 0] - 0.25) ** 2 + (x[:, 1] - 0.25) ** 2) / (0.2 ** 2))


def run(epochs):
    torch.manual_seed(0)
    NF = PWQuadManager(n_flow=2)
    NF.create_model(2, 4, [3] * 3)
    optim = torch.optim.Adamax(NF._model.parameters(), lr=2e-3, weight_decay=1e-04)
    torch.cuda.synchronize()
    t0 = time.time()
    NF._train_variance_forward_seq(camel, optim, True, tempfile.mkdtemp(), 10000, epochs, 0, False, True, preburn_time=50)
    torch.cuda.synchronize()
    return time.time() - t0


run(5)
print("60 epochs: %.3f s" % run(60))
pr = cProfile.Profile()
pr.enable()
run(60)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run(10)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30))
