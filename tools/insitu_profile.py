#!/usr/bin/env python
"""In-situ kernel durations of one cfg2 train-mode forward at 2^22 points (torch.profiler / CUPTI: warm caches,
back-to-back launches) — the counterpart of the serialised, cold-cache ncu launch list."""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from nf_b200.normalizing_flows.manager import PWLinManager  # noqa: E402

torch.manual_seed(1234)
NF = PWLinManager(n_flow=8)
NF.create_model(4, 6, 32, [64] * 3, 4)
model = NF._model.train()
x = torch.rand(1 << 22, 8, device="cuda", dtype=torch.float32)
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
agg = collections.OrderedDict()
for e in prof.events():
    if e.device_type is not None and "cuda" in str(e.device_type).lower() and e.device_time > 0:
        a = agg.setdefault(e.name[:70], [0, 0.0])
        a[0] += 1
        a[1] += e.device_time
tot = sum(a[1] for a in agg.values())
print("total kernel time per forward: %.3f ms" % (tot / 3 / 1e3))
for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:8]:
    print("%-72s n/fwd=%5.1f  avg=%8.1f us  share=%5.1f%%" % (n, a[0] / 3, a[1] / a[0], 100 * a[1] / tot))

seq = [e for e in prof.events() if e.device_time > 0 and ("flow_" in e.name)]
seq.sort(key=lambda e: e.time_range.start)
per = len(seq) // 3
print("launch sequence of one forward (us): " + " ".join("%s%.0f" % ("M" if "moments" in e.name else ("p" if "pack" in e.name else "T"), e.device_time) for e in seq[:per]))
