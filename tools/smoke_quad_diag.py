"""Diagnostic: distribution of the PWQuad forward error of smoke()'s case (h kernel vs tf32 kernel vs oracle)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import flow as oflow                                   # noqa: E402
from nf_b200.normalizing_flows.manager import PWQuadManager       # noqa: E402

torch.manual_seed(2)
NQ = PWQuadManager(n_flow=8)
NQ.create_model(6, 32, [64] * 3)
mq = NQ._model.train()
sdq = {k: (v.detach().double().cpu() if v.dtype.is_floating_point else v.cpu()) for k, v in mq.state_dict().items()}
xq = torch.rand(2048, 8, dtype=torch.float32).double()
xjq = NQ.format_input(xq, torch.device("cuda:0"))
with torch.no_grad():
    XQ = mq(xjq)
    refq, _ = oflow.flow_forward(oflow.pwquad_layers(8, 6), sdq, xjq.cpu(), "quad", 32, train=True)
    ref32, _ = oflow.flow_forward(oflow.pwquad_layers(8, 6), {k: (v.float() if v.dtype.is_floating_point else v) for k, v in sdq.items()},
                                  xjq.cpu().float(), "quad", 32, train=True)
for name, X in (("gpu", XQ.cpu().double()), ("oracle fp32", ref32.double())):
    lj = (torch.log(X[:, -1]) - torch.log(refq[:, -1])).abs()
    pe = (X[:, :-1] - refq[:, :-1]).abs().max(1).values
    rl = lj / torch.log(refq[:, -1]).abs().clamp_min(1.0)
    print(name, "rel q99.9 %.2e max %.2e" % (float(torch.quantile(rl, 0.999)), float(rl.max())))
    print(name, "logJ: median %.2e q99 %.2e q99.9 %.2e max %.2e | points: q99.9 %.2e max %.2e | max |log J| %.2f" % (
        float(lj.median()), float(torch.quantile(lj, 0.99)), float(torch.quantile(lj, 0.999)), float(lj.max()),
        float(torch.quantile(pe, 0.999)), float(pe.max()), float(torch.log(refq[:, -1]).abs().max())))
    top = torch.topk(lj, 5)
    print("   worst rows", top.indices.tolist(), ["%.2e" % v for v in top.values.tolist()],
          "their point errors", ["%.2e" % float(pe[i]) for i in top.indices])
