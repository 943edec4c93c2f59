#!/usr/bin/env python
"""Times the fused RAMBO kernel (cfg3: 2->4 massive, all cuts) on 2^24 events.  Development aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nf_b200.PhaseSpace.flat_phase_space_generator import FlatInvertiblePhasespace  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
ps = FlatInvertiblePhasespace([100.0] * 2, [100.0] * 4)
ps.check_nan = False
r = torch.rand(n, 8, device="cuda", dtype=torch.float64)
for name, kw in (("momenta+weight, all cuts", dict(pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5)),
                 ("weight only, all cuts", dict(pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5, momenta=False)),
                 ("momenta+weight, no cuts", dict())):
    for _ in range(3):
        ps.generateKinematics_batch(1000.0, r, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ps.generateKinematics_batch(1000.0, r, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-28s %.3f ms  %.3e events/s  %.0f GB/s (264 B/event)" % (name, ms, n / ms * 1e3, n * 264 / ms / 1e6))
