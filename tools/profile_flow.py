#!/usr/bin/env python
"""Small driver for ncu: runs the cfg2 flow forward (eval or train BN) a few times on 2^22 points, or the
RAMBO kernel on 2^24 events.  Usage: python tools/profile_flow.py {eval|train|rambo} [log2_points]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "eval"
if mode == "rambo":
    from nf_b200.PhaseSpace.flat_phase_space_generator import FlatInvertiblePhasespace
    n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 24)
    ps = FlatInvertiblePhasespace([100.0] * 2, [100.0] * 4)
    ps.check_nan = False
    r = torch.rand(n, 8, device="cuda", dtype=torch.float64)
    for _ in range(3):
        ps.generateKinematics_batch(1000.0, r, pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5)
    torch.cuda.synchronize()
else:
    from nf_b200.normalizing_flows.manager import PWLinManager
    n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 22)
    torch.manual_seed(1234)
    NF = PWLinManager(n_flow=8)
    NF.create_model(4, 6, 32, [64] * 3, 4)
    model = NF._model.train(mode == "train")
    x = torch.rand(n, 8, device="cuda", dtype=torch.float32)
    with torch.no_grad():
        for _ in range(2):
            model(x)
    torch.cuda.synchronize()
print("done", mode)
