import os, sys, time, tempfile
sys.path.insert(0, "/root/repo")
import torch
from nf_b200.normalizing_flows.manager import PWQuadManager
def camel(x):
    return torch.exp(-((x[:, 0] - 0.75) ** 2 + (x[:, 1] - 0.75) ** 2) / (0.2 ** 2)) + \
        torch.exp(-((x[:, 0] - 0.25) ** 2 + (x[:, 1] - 0.25) ** 2) / (0.2 ** 2))
ep = int(sys.argv[1])
torch.manual_seed(0)
NF = PWQuadManager(n_flow=2); NF.create_model(2, 4, [3] * 3)
optim = torch.optim.Adamax(NF._model.parameters(), lr=2e-3, weight_decay=1e-04)
torch.cuda.synchronize(); t0 = time.time()
NF._train_variance_forward_seq(camel, optim, True, tempfile.mkdtemp(), 10000, ep, 0, False, True, preburn_time=50)
torch.cuda.synchronize(); print("epochs", ep, "streams", os.environ.get("NIS_TRAIN_STREAMS", "8"), "graph", os.environ.get("NIS_TRAIN_GRAPH","1"), "%.2f s" % (time.time() - t0))
