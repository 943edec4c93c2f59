#!/usr/bin/env python
"""Development aid: the README example's wall time inside the bench process, after each of the other extras."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
which = sys.argv[1] if len(sys.argv) > 1 else "none"
if which == "wide":
    bench.bench_wide(3, 3, 1, dev)
elif which == "integrate":
    bench.bench_integrate(1, dev)
elif which == "train":
    bench.bench_train_step(3, 3, 1, dev)
    bench.bench_train_step(3, 3, 1, dev, log2n=20)
elif which == "parity":
    bench.parity_checks(0, 1, dev)
for rep in range(2):
    r = bench.bench_readme(dev)
    print("after %-9s rep %d: %.2f s  best_loss %.4f" % (which, rep, r["value"], r["best_loss"]))
