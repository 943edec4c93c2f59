"""Helpers shared by the GPU parity tests (everything goes through the public API -> C ABI)."""
import numpy as np
import torch

from oracle import flow as oflow

from nf_b200.normalizing_flows.manager import PWLinManager, PWQuadManager

RTOL, ATOL_Y = 1e-5, 1e-6          # north_star: 1e-5 relative on points and log-Jacobians, in fp32


def make_manager(meta):
    if meta["kind"] == "quad":
        NF = PWQuadManager(n_flow=meta["n_flow"])
        NF.create_model(meta["n_cells"], meta["n_bins"], meta["NN"])
    else:
        NF = PWLinManager(n_flow=meta["n_flow"])
        NF.create_model(meta["n_pass_through"], meta["n_cells"], meta["n_bins"], meta["NN"], meta["roll_step"])
    return NF


def oracle_layers(meta):
    if meta["kind"] == "quad":
        return oflow.pwquad_layers(meta["n_flow"], meta["n_cells"])
    return oflow.pwlin_layers(meta["n_flow"], meta["n_pass_through"], meta["n_cells"], meta["roll_step"])


def compare_flow(XJ, bins, ref_XJ, ref_bins, what=""):
    """XJ [B,d+1] (ours, any float dtype, cpu), bins [C,B,d] int32 (ours); ref_bins: list of [B,T_c].
    Bin indices must be identical except where fp32 rounding puts x on the other side of an edge
    (|delta| == 1, at most a handful); points and log-Jacobians within 1e-5 relative."""
    XJ = XJ.double()
    B = XJ.shape[0]
    flipped = np.zeros(B, bool)
    nflip = 0
    total = 0
    for c, rb in enumerate(ref_bins):
        rb = np.asarray(rb)
        ob = bins[c, :, :rb.shape[1]].numpy()
        diff = ob != rb
        assert np.all(np.abs(ob - rb)[diff] == 1), "%s: bin off by more than one in cell %d" % (what, c)
        flipped |= diff.any(1)
        nflip += int(diff.sum())
        total += diff.size
    assert nflip <= max(2, int(2e-4 * total)), "%s: %d of %d bins differ" % (what, nflip, total)
    keep = ~flipped
    y, ry = XJ[keep, :-1], ref_XJ[keep, :-1]
    assert torch.allclose(y, ry, rtol=RTOL, atol=ATOL_Y), "%s: points, max abs err %g" % (what, float((y - ry).abs().max()))
    lj, rlj = torch.log(XJ[keep, -1]), torch.log(ref_XJ[keep, -1])
    err = (lj - rlj).abs() / rlj.abs().clamp_min(1.0)
    assert float(err.max()) <= RTOL, "%s: log-Jacobian rel err %g" % (what, float(err.max()))
    return nflip
