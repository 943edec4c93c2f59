"""Helpers shared by the GPU parity tests (everything goes through the public API -> C ABI)."""
import numpy as np
import torch

from oracle import flow as oflow

from nf_b200.normalizing_flows.manager import AffineManager, PWLinManager, PWQuadManager

RTOL, ATOL_Y = 1e-5, 1e-6          # north_star: 1e-5 relative on points and log-Jacobians, in fp32


def make_manager(meta):
    if meta["kind"] == "quad":
        NF = PWQuadManager(n_flow=meta["n_flow"])
        NF.create_model(meta["n_cells"], meta["n_bins"], meta["NN"])
    elif meta["kind"] == "affine":
        NF = AffineManager(n_flow=meta["n_flow"])
        NF.create_model(meta["n_pass_through"], meta["n_cells"], meta["NN"], meta["roll_step"])
    else:
        NF = PWLinManager(n_flow=meta["n_flow"])
        NF.create_model(meta["n_pass_through"], meta["n_cells"], meta["n_bins"], meta["NN"], meta["roll_step"])
    return NF


def oracle_layers(meta):
    if meta["kind"] == "quad":
        return oflow.pwquad_layers(meta["n_flow"], meta["n_cells"])
    return oflow.pwlin_layers(meta["n_flow"], meta["n_pass_through"], meta["n_cells"], meta["roll_step"])


EDGE_TOL = 2e-6                    # a bin may differ from the float64 reference only for a coordinate this close to
                                   # a bin edge (the parity tolerance on that coordinate: ATOL_Y + RTOL * x <= 1.1e-5,
                                   # bound used: 2e-6 + 1e-5 * distance scale below) - every flip is checked
FLIP_BUDGET = 1e-4                 # and at most this fraction of all bins (n_bins * 2 * allowed coordinate error)


COND_K = 8.0                       # per-point log-Jacobian bound: RTOL + COND_K * 2^-24 * (condition of the point), below


def compare_flow(XJ, bins, ref_XJ, ref_bins, what="", fp32_yardstick=None, ref_edges=None, max_log_j=None, ref_cond=None):
    """XJ [B,d+1] (ours, any float dtype, cpu), bins [C,B,d] int32 (ours); ref_bins: list of [B,T_c].
    Bin indices must be identical to the float64 reference except where float32 rounding puts a coordinate on the
    other side of an edge: |delta| == 1 and — when the oracle's edge distances ``ref_edges`` (list of [B,T_c]) are
    given — the float64 coordinate provably within EDGE_TOL of that edge; the count is printed and bounded by
    FLIP_BUDGET.  Points and log-Jacobians within 1e-5 relative.

    ``fp32_yardstick`` (the oracle itself evaluated in float32, [B,d+1]) switches the log-Jacobian
    check to the large-batch form: the Jacobian of a point inside a narrow bin is ill-conditioned
    (d log f / d logit ~ 1/W_k), so over 10^5 spline evaluations the worst point of ANY float32
    evaluation exceeds 1e-5; there we require the 99.9 % quantile <= 1e-5 and the maximum to be no
    worse than twice what the float32 oracle itself loses (and the median <= 3e-6); ``max_log_j`` is an absolute
    cap on the maximum on top of that.

    ``ref_cond`` (PWQuad: the oracle's per-cell sensitivities sum_t |d log f_t / d x_t|, list of [B]) replaces the
    rule on the maximum - the largest of 10^5 heavy-tailed errors is a lottery that any two float32 evaluations
    win or lose by a factor of a few - by a bound on EVERY point: err <= RTOL + COND_K * 2^-24 * condition, i.e. the
    float32 rounding of the coordinates a cell receives, times what the narrow bin multiplies it by (measured on
    2.6e5 points of cfg4: this kernel reaches 2.6 x 2^-24 x condition, the float32 oracle 3.7)."""
    XJ = XJ.double()
    B = XJ.shape[0]
    flipped = np.zeros(B, bool)
    nflip = 0
    total = 0
    worst_edge = 0.0
    for c, rb in enumerate(ref_bins):
        rb = np.asarray(rb)
        ob = bins[c, :, :rb.shape[1]].numpy()
        diff = ob != rb
        assert np.all(np.abs(ob - rb)[diff] == 1), "%s: bin off by more than one in cell %d" % (what, c)
        if ref_edges is not None and diff.any():
            dist = np.asarray(ref_edges[c])[diff]
            worst_edge = max(worst_edge, float(dist.max()))
            assert float(dist.max()) <= EDGE_TOL, \
                "%s: cell %d: a bin differs although the coordinate is %g away from the edge" % (what, c, dist.max())
        flipped |= diff.any(1)
        nflip += int(diff.sum())
        total += diff.size
    print("%s: %d of %d bins differ from the float64 reference (%.2e); farthest from its edge: %.2e" % (
        what, nflip, total, nflip / max(total, 1), worst_edge))
    assert nflip <= max(2, int(FLIP_BUDGET * total)), "%s: %d of %d bins differ" % (what, nflip, total)
    keep = ~flipped
    y, ry = XJ[keep, :-1], ref_XJ[keep, :-1]
    yn = ((y - ry).abs() / (ATOL_Y + RTOL * ry.abs())).reshape(-1)         # error in units of the tolerance
    if fp32_yardstick is None:
        assert float(yn.max()) <= 1.0, "%s: points, max abs err %g" % (what, float((y - ry).abs().max()))
    else:
        # large batches: like the log-Jacobian below, the worst of ~10^7 float32 spline evaluations sits in a narrow bin
        yy = ((fp32_yardstick.double()[keep, :-1] - ry).abs() / (ATOL_Y + RTOL * ry.abs())).reshape(-1)
        yy = yy[torch.isfinite(yy)]
        k = max(1, int(1e-4 * yn.numel()))
        q = float(torch.topk(yn, k).values[-1]) if yn.numel() > 1 else float(yn.max())
        msg = "%s: points, error / tolerance: ours q99.99 %.2f max %.2f | float32 oracle max %.2f" % (
            what, q, float(yn.max()), float(yy.max()))
        print(msg)
        assert q <= 1.0 and float(yn.max()) <= max(1.0, 2 * float(yy.max())), msg
    lj, rlj = torch.log(XJ[keep, -1]), torch.log(ref_XJ[keep, -1])
    err = (lj - rlj).abs() / rlj.abs().clamp_min(1.0)
    if fp32_yardstick is None:
        assert float(err.max()) <= RTOL, "%s: log-Jacobian rel err %g" % (what, float(err.max()))
    else:
        ylj = torch.log(fp32_yardstick.double()[keep, -1])
        yerr = (ylj - rlj).abs() / rlj.abs().clamp_min(1.0)
        yerr = yerr[torch.isfinite(yerr)]
        q999, yq999 = float(torch.quantile(err, 0.999)), float(torch.quantile(yerr, 0.999))
        msg = "%s: log-Jacobian rel err: ours median %.2e q99.9 %.2e max %.2e | float32 oracle median %.2e q99.9 %.2e max %.2e" % (
            what, float(err.median()), q999, float(err.max()), float(yerr.median()), yq999, float(yerr.max()))
        print(msg)
        assert float(err.median()) <= 0.3 * RTOL, msg
        assert q999 <= max(RTOL, 2 * yq999), msg
        if ref_cond:
            cond = torch.stack([torch.as_tensor(np.asarray(c), dtype=torch.float64) for c in ref_cond]).sum(0)[keep]
            bound = RTOL + COND_K * 2.0 ** -24 * cond / rlj.abs().clamp_min(1.0)
            worst = float((err / bound).max())
            print("%s: log-Jacobian error / (RTOL + %g x 2^-24 x condition): max %.2f" % (what, COND_K, worst))
            assert worst <= 1.0, msg
        else:
            assert float(err.max()) <= max(RTOL, 2 * float(yerr.max())), msg
        if max_log_j is not None:
            assert float(err.max()) <= max_log_j, msg
    return nflip
