#!/usr/bin/env python
"""Tail of the forward error against the float64 oracle, next to the float32 oracle's own (development aid: A/B of
two library builds through NIS_LIB_PATH).  usage: parity_tail.py [log2 batch] [seeds]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402
from oracle import flow as oflow  # noqa: E402
from gpu_util import make_manager, oracle_layers  # noqa: E402

B = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 16)
seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 2
CFGS = [dict(name="cfg4", kind="quad", n_flow=8, n_cells=6, n_bins=32, NN=[64] * 3),
        dict(name="cfg2", kind="lin", n_flow=8, n_pass_through=4, n_cells=6, n_bins=32, NN=[64] * 3, roll_step=4)]


def stats(err):
    return "median %.2e q99.9 %.2e q99.99 %.2e max %.2e  >1e-5: %d  >1e-4: %d" % (
        float(err.median()), float(torch.quantile(err, 0.999)), float(torch.topk(err, max(1, err.numel() // 10000)).values[-1]),
        float(err.max()), int((err > 1e-5).sum()), int((err > 1e-4).sum()))


for cfg in CFGS:
    for mode in ("train", "eval"):
        ours, yard, ro, ry = [], [], [], []
        for seed in range(seeds):
            torch.manual_seed(5 + seed)
            NF = make_manager(cfg)
            cells, _ = oflow.compile_layers(oracle_layers(cfg), cfg["n_flow"])
            sd = oflow.init_state_dict(cells, cfg["n_flow"], cfg["kind"], cfg["n_bins"], cfg["NN"], seed=7 + seed,
                                       dtype=torch.float32, bn_jitter=0.2)
            NF._model.load_state_dict(sd)
            NF._model.train(mode == "train")
            x = torch.rand(B, cfg["n_flow"], generator=torch.Generator().manual_seed(2026 + seed), dtype=torch.float32).double()
            xj = NF.format_input(x, torch.device("cuda"))
            with torch.no_grad():
                XJ = NF._model(xj).cpu().double()
                sd64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in sd.items()}
                cond = []
                ref, _ = oflow.flow_forward(oracle_layers(cfg), sd64, xj.cpu(), cfg["kind"], cfg["n_bins"],
                                            train=(mode == "train"), clamp_bins=True, cond=cond)
                ref32, _ = oflow.flow_forward(oracle_layers(cfg), sd, xj.cpu().float(), cfg["kind"], cfg["n_bins"],
                                              train=(mode == "train"), clamp_bins=True)
            rl = torch.log(ref[:, -1])
            ours.append((torch.log(XJ[:, -1]) - rl).abs() / rl.abs().clamp_min(1.0))
            e32 = (torch.log(ref32[:, -1].double()) - rl).abs() / rl.abs().clamp_min(1.0)
            yard.append(e32[torch.isfinite(e32)])
            if cond:
                ct = torch.stack(cond).sum(0) / rl.abs().clamp_min(1.0)
                ro.append((ours[-1] - 1e-5).clamp_min(0) / (6e-8 * ct))
                ry.append(((e32 - 1e-5).clamp_min(0) / (6e-8 * ct))[torch.isfinite(e32)])
        print("%s/%s B=%d x %d  ours:       %s" % (cfg["name"], mode, B, seeds, stats(torch.cat(ours))))
        print("%s/%s B=%d x %d  f32 oracle: %s" % (cfg["name"], mode, B, seeds, stats(torch.cat(yard))))
        if ro:
            print("%s/%s (err - 1e-5) / (eps32 x condition): ours max %.2f  f32 oracle max %.2f" % (
                cfg["name"], mode, float(torch.cat(ro).max()), float(torch.cat(ry).max())))
