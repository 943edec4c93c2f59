#!/usr/bin/env python
"""Distributed run of the training loop itself (torchrun, one rank per GPU): the README integrand with 6 minibatches per
epoch (3 per rank at 2 ranks, so every rank also runs its minibatches concurrently), a few epochs; afterwards every rank
must hold the same finite weights and BatchNorm buffers must be finite.  Development aid / multi-GPU check."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from nf_b200.normalizing_flows.manager import PWQuadManager  # noqa: E402


def camel(x):
    return torch.exp(-((x[:, 0] - 0.75) ** 2 + (x[:, 1] - 0.75) ** 2) / (0.2 ** 2)) + \
        torch.exp(-((x[:, 0] - 0.25) ** 2 + (x[:, 1] - 0.25) ** 2) / (0.2 ** 2))


rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(0)
NF = PWQuadManager(n_flow=2)
NF.create_model(2, 4, [3] * 3, dev=local)
optim = torch.optim.Adamax(NF._model.parameters(), lr=2e-3, weight_decay=1e-04)
NF._train_variance_forward_seq(camel, optim, False, tempfile.mkdtemp(), 12000, 12, 0, False, True, dev=local, preburn_time=4)
flat = torch.cat([p.detach().reshape(-1).float() for p in NF._model.parameters()])
bufs = torch.cat([b.detach().reshape(-1).float() for b in NF._model.buffers() if b.dtype.is_floating_point])
allp = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(allp, flat)
same = all(torch.equal(allp[0], a) for a in allp[1:])
fin = bool(torch.isfinite(flat).all()) and bool(torch.isfinite(bufs).all())
sig, err = NF.integrate(camel, 4, 20000, local)
if rank == 0:
    print("world %d: weights identical on every rank: %s, finite: %s, best_loss %.4f from %.4f, integrate %.5f +- %.5f (0.23232)"
          % (world, same, fin, float(NF.best_loss), float(NF.int_loss), float(sig), float(err)))
    assert same and fin and float(NF.best_loss) < float(NF.int_loss)
dist.destroy_process_group()
