#!/usr/bin/env python
"""Wall time of the reference's README example (BASELINE configs[0]): 2-D camel, PWQuadManager, 4 bins, MLP [3]*3,
10000 points per batch, 300 epochs of Adamax variance training, then integrate.  Development aid."""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nf_b200.normalizing_flows.manager import PWQuadManager  # noqa: E402


def camel(x):
    return torch.exp(-((x[:, 0] - 0.75) ** 2 + (x[:, 1] - 0.75) ** 2) / (0.2 ** 2)) + \
        torch.exp(-((x[:, 0] - 0.25) ** 2 + (x[:, 1] - 0.25) ** 2) / (0.2 ** 2))


for rep in range(3):          # the first repetition pays for lazy module loading and graph capture on a cold process
    torch.manual_seed(0)
    NF = PWQuadManager(n_flow=2)
    NF.create_model(2, 4, [3] * 3)
    optim = torch.optim.Adamax(NF._model.parameters(), lr=2e-3, weight_decay=1e-04)
    torch.cuda.synchronize()
    t0 = time.time()
    NF._train_variance_forward_seq(camel, optim, True, tempfile.mkdtemp(), 10000, 300, 0, False, True, preburn_time=50)
    torch.cuda.synchronize()
    t1 = time.time()
    sig, err = NF.integrate(camel, 10, 10000, 0)
    torch.cuda.synchronize()
    t2 = time.time()
    print("README example (rep %d): train 300 epochs %.2f s (%.2f ms/epoch), integrate %.3f s -> %.5f +- %.5f (analytic 0.23232); "
          "best_loss %.4f from int_loss %.4f" % (rep, t1 - t0, (t1 - t0) / 300 * 1e3, t2 - t1, float(sig), float(err),
                                                 float(NF.best_loss), float(NF.int_loss)))
