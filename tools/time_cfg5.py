#!/usr/bin/env python
"""Times BASELINE configs[4] (16-D PWQuad flow, 64 bins, MLP [256]*4): forward (eval / train-mode BN) and one
variance-loss training step on 2^n points.  Development aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nf_b200.normalizing_flows.manager import PWQuadManager  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 16)
torch.manual_seed(1234)
NF = PWQuadManager(n_flow=16)
NF.create_model(8, 64, [256] * 4)
x = torch.rand(n, 16, device="cuda", dtype=torch.float32)
f = torch.exp(-((x - 0.5) ** 2).sum(-1) / 0.4)
FLOP = 8 * 2 * 462848


def timed(fn, reps=3):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for mode in ("eval", "train"):
    model = NF._model.train(mode == "train")
    with torch.no_grad():
        ms = timed(lambda: model(x))
    print("cfg5 fwd %s B=%d: %.3f ms  %.3e pts/s  %.1f TFLOP/s algorithmic" % (mode, n, ms, n / ms * 1e3, n * FLOP / ms / 1e9))
model = NF._model.train()


def step():
    model.zero_grad()
    XJ = model(x)
    torch.var(f * XJ[:, -1]).backward()


try:
    ms = timed(step)
    print("cfg5 fwd+bwd B=%d: %.3f ms  %.3e pts/s" % (n, ms, n / ms * 1e3))
except Exception as e:  # noqa: BLE001
    print("cfg5 fwd+bwd failed:", e)
