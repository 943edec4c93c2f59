"""GPU parity of the fused backward (autograd.Function -> nis_flow_backward) against the reference's
autograd gradients (golden) and torch.autograd of the oracle at larger sizes.

Tolerance: gradients are float32 sums over the batch; each tensor must match the float64 reference to
2e-4 of that tensor's largest entry plus 1e-5 of the largest gradient entry of the whole model (some
gradients are identically zero by symmetry — e.g. a BatchNorm shift that the next train-mode BatchNorm
removes — and only rounding noise is left there)."""
import pytest
import torch

from conftest import GRAD_CASES
from gpu_util import make_manager, oracle_layers
from oracle import flow as oflow
from oracle import nis as onis

pytestmark = pytest.mark.gpu
GTOL = 2e-4


def close(a, ref, name, gscale, flips=False):
    """flips=True (large batches): the parameter gradient is discontinuous where a point sits exactly on a
    ReLU kink or a bin edge; over ~10^7 unit evaluations a handful of points land within float32 rounding
    of one and take the other branch than the float64 oracle, which moves one row of one weight gradient
    by O(1/sqrt(B)) of its size.  The bulk (95 % quantile of the entries) must still meet the tight bound;
    the maximum gets a loose sanity bound."""
    scale = float(ref.abs().max())
    diff = (a.double().cpu() - ref).abs().reshape(-1)
    tight = GTOL * scale + 1e-5 * gscale
    if not flips:
        assert float(diff.max()) <= tight, "%s: err %g vs scale %g (model scale %g)" % (name, float(diff.max()), scale, gscale)
    else:
        q = float(torch.quantile(diff, 0.95)) if diff.numel() > 20 else float(diff.max())
        assert q <= tight or float(diff.max()) <= 5 * tight, "%s: q95 err %g vs scale %g (model scale %g)" % (name, q, scale, gscale)
        assert float(diff.max()) <= 2e-2 * gscale, "%s: max err %g (model scale %g)" % (name, float(diff.max()), gscale)


@pytest.mark.parametrize("case", GRAD_CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_variance_loss_gradients_match_reference(golden, case, mode):
    g = golden("flow_" + case)
    NF = make_manager(g.meta)
    model = NF._model
    model.load_state_dict(g.state_dict())
    model.train(mode == "train")
    xj = g.t("xj").cuda()
    model.zero_grad()
    XJ = model(xj)
    fres = g.t(mode + "/grad_var/fres").cuda()
    loss = torch.var(fres * XJ[:, -1] / fres.max())             # manager.py:245-255
    assert abs(float(loss) - float(g[mode + "/grad_var/loss"])) <= 1e-5 * abs(float(g[mode + "/grad_var/loss"]))
    loss.backward()
    gscale = max(float(g.t("%s/grad_var/%s" % (mode, k)).abs().max()) for k, _ in model.named_parameters())
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        close(p.grad, g.t("%s/grad_var/%s" % (mode, k)), k, gscale)


@pytest.mark.parametrize("case", GRAD_CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
@pytest.mark.parametrize("io", [torch.float64, torch.float32])
def test_generic_upstream_gradient_matches_reference(golden, case, mode, io):
    g = golden("flow_" + case)
    NF = make_manager(g.meta)
    model = NF._model
    model.load_state_dict(g.state_dict())
    model.train(mode == "train")
    xin = g.t("xj").to(io).cuda().requires_grad_(True)
    G = g.t(mode + "/grad_lin/G").to(io).cuda()
    model.zero_grad()
    (model(xin) * G).sum().backward()
    assert xin.grad.dtype == io
    gscale = max(float(g.t("%s/grad_lin/%s" % (mode, k)).abs().max()) for k, _ in model.named_parameters())
    close(xin.grad, g.t(mode + "/grad_lin/dxj"), "dxj", float(g.t(mode + "/grad_lin/dxj").abs().max()))
    for k, p in model.named_parameters():
        close(p.grad, g.t("%s/grad_lin/%s" % (mode, k)), k, gscale)


BIG = [
    dict(name="cfg2", kind="lin", n_flow=8, n_pass_through=4, n_cells=6, n_bins=32, NN=[64] * 3, roll_step=4, B=5000),
    dict(name="cfg4", kind="quad", n_flow=8, n_cells=6, n_bins=32, NN=[64] * 3, B=3000),
    dict(name="cfg1", kind="quad", n_flow=2, n_cells=2, n_bins=4, NN=[3] * 3, B=2000),
    dict(name="cfg5_small", kind="quad", n_flow=16, n_cells=8, n_bins=16, NN=[128] * 2, B=1500),
    # AffineCoupling: biased hidden layers folded into the running mean, Reshape(2, T) output rows
    dict(name="affine8d", kind="affine", n_flow=8, n_pass_through=4, n_cells=4, n_bins=1, NN=[64] * 2, roll_step=4, B=3000),
]


@pytest.mark.parametrize("cfg", BIG, ids=[c["name"] for c in BIG])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_gradients_match_oracle_autograd_at_size(cfg, mode):
    torch.manual_seed(3)
    NF = make_manager(cfg)
    model = NF._model
    layers = oracle_layers(cfg)
    cells, _ = oflow.compile_layers(layers, cfg["n_flow"])
    sd = oflow.init_state_dict(cells, cfg["n_flow"], cfg["kind"], cfg["n_bins"], cfg["NN"], seed=9,
                               dtype=torch.float32, bn_jitter=0.2)
    model.load_state_dict(sd)
    model.train(mode == "train")
    gen = torch.Generator().manual_seed(77)
    x = torch.rand(cfg["B"], cfg["n_flow"], generator=gen, dtype=torch.float32).double()
    fres = torch.exp(-((x - 0.5) ** 2).sum(-1) / 0.2)
    # two minibatches accumulated before one backward, like the training loop (manager.py:219-278)
    halves = [slice(0, cfg["B"] // 2), slice(cfg["B"] // 2, cfg["B"])]
    model.zero_grad()
    loss = 0
    for h in halves:
        XJ = model(NF.format_input(x[h], torch.device("cuda")))
        loss = loss + torch.var(fres[h].cuda() * XJ[:, -1] / fres.max())
    (loss / 2).backward()
    sd64 = {k: (v.double().requires_grad_(k.endswith("weight") or k.endswith("bias")) if v.dtype.is_floating_point else v)
            for k, v in sd.items()}
    ref = 0
    for h in halves:
        xj = torch.cat((x[h], torch.ones(x[h].shape[0], 1, dtype=torch.float64)), 1)
        XJr, _ = oflow.flow_forward(layers, sd64, xj, cfg["kind"], cfg["n_bins"], train=(mode == "train"))
        ref = ref + onis.minibatch_loss(fres[h], XJr[:, -1], fres.max(), "var")
    (ref / 2).backward()
    assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref))
    gscale = max(float(sd64[k].grad.abs().max()) for k, _ in model.named_parameters())
    for k, p in model.named_parameters():
        close(p.grad, sd64[k].grad, k, gscale, flips=True)


@pytest.mark.parametrize("B", [2048, (1 << 13) + 77])
def test_tensor_core_backward_agrees_with_generic_backward_and_oracle(monkeypatch, B):
    """Train-mode backward of cfg2 runs on the tcgen05 kernels (flow_bwd_tc.cu) by default; NIS_BWD_TC=0 selects
    the shape-generic kernel.  Both must meet the parity bar against float64 autograd through the oracle, for the
    parameter gradients and for the gradient with respect to the input points (ragged last tile included)."""
    cfg = dict(BIG[0], B=B)
    layers = oracle_layers(cfg)
    cells, _ = oflow.compile_layers(layers, cfg["n_flow"])
    sd = oflow.init_state_dict(cells, cfg["n_flow"], cfg["kind"], cfg["n_bins"], cfg["NN"], seed=21,
                               dtype=torch.float32, bn_jitter=0.2)
    gen = torch.Generator().manual_seed(5)
    x = torch.rand(B, cfg["n_flow"], generator=gen, dtype=torch.float32).double()
    fres = torch.exp(-((x - 0.5) ** 2).sum(-1) / 0.2)
    sd64 = {k: (v.double().requires_grad_(k.endswith("weight") or k.endswith("bias")) if v.dtype.is_floating_point else v)
            for k, v in sd.items()}
    xj64 = torch.cat((x, torch.ones(B, 1, dtype=torch.float64)), 1).requires_grad_(True)
    XJr, _ = oflow.flow_forward(layers, sd64, xj64, cfg["kind"], cfg["n_bins"], train=True)
    ref = onis.minibatch_loss(fres, XJr[:, -1], fres.max(), "var") + (XJr[:, :-1] ** 2).mean()
    ref.backward()
    gscale = max(float(v.grad.abs().max()) for v in sd64.values() if getattr(v, "grad", None) is not None)
    grads = {}
    for backend, env in (("tcgen05", None), ("generic", "0")):
        monkeypatch.delenv("NIS_BWD_TC", raising=False)
        if env is not None:
            monkeypatch.setenv("NIS_BWD_TC", env)
        torch.manual_seed(3)
        NF = make_manager(cfg)
        model = NF._model
        model.load_state_dict(sd)
        model.train()
        xj = NF.format_input(x, torch.device("cuda")).requires_grad_(True)
        XJ = model(xj)
        loss = torch.var(fres.cuda() * XJ[:, -1] / fres.max()) + (XJ[:, :-1] ** 2).mean()
        loss.backward()
        assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref))
        for k, p in model.named_parameters():
            close(p.grad, sd64[k].grad, backend + ":" + k, gscale, flips=True)
        close(xj.grad, xj64.grad, backend + ":dL/dx", float(xj64.grad.abs().max()), flips=True)
        grads[backend] = {k: p.grad.detach().cpu().double() for k, p in model.named_parameters()}
    for k in grads["tcgen05"]:
        close(grads["tcgen05"][k], grads["generic"][k], "tc vs generic:" + k, gscale, flips=True)


@pytest.mark.parametrize("backend", ["tcgen05", "generic"])
def test_full_size_cfg5_gradients(monkeypatch, backend):
    """BASELINE configs[4] at its full shape (16-D PWQuad, 8 cells, 64 bins, MLP [256]*4), gradients against
    float64 autograd through the oracle: by default the streamed-weights tcgen05 backward (flow_bwd_wide.cu);
    with NIS_BWD_TC=0 the shape-generic kernel, whose train-mode launches keep the activations of their one
    step in three rotating shared-memory buffers (which is what makes this width fit)."""
    monkeypatch.delenv("NIS_BWD_TC", raising=False)
    if backend == "generic":
        monkeypatch.setenv("NIS_BWD_TC", "0")
    cfg = dict(name="cfg5", kind="quad", n_flow=16, n_cells=8, n_bins=64, NN=[256] * 4, B=2600)
    test_gradients_match_oracle_autograd_at_size(cfg, "train")


@pytest.mark.parametrize("which,log2b", [(0, 19), (3, 15), (3, 18)], ids=["cfg2_2p19", "cfg5_small_2p15", "cfg5_small_2p18"])
def test_tensor_core_backward_at_large_batch_matches_generic(monkeypatch, which, log2b):
    """Many tiles per CTA (cfg2: 2^19 points = 28 tiles per CTA; the 128-wide cfg5_small: 2^15 points, and 2^18 points =
    14 tiles per hidden-layer CTA / 42 per output-layer CTA of the streamed-weights wgrad kernel, so that its second and
    later accumulator flushes -- vector reductions onto the slice -- run): the weight-gradient accumulators in tensor
    memory are flushed to the CTA's slice every 16 (8) tiles because tcgen05 accumulation truncates; the result must
    agree with the FP32 generic kernel (itself pinned against the oracle above)."""
    cfg = dict(which if isinstance(which, dict) else BIG[which], B=1 << log2b)
    layers = oracle_layers(cfg)
    cells, _ = oflow.compile_layers(layers, cfg["n_flow"])
    sd = oflow.init_state_dict(cells, cfg["n_flow"], cfg["kind"], cfg["n_bins"], cfg["NN"], seed=33,
                               dtype=torch.float32, bn_jitter=0.2)
    gen = torch.Generator().manual_seed(6)
    x = torch.rand(cfg["B"], cfg["n_flow"], generator=gen, dtype=torch.float32)
    fres = torch.exp(-((x - 0.5) ** 2).sum(-1) / 0.2).cuda()
    grads = {}
    for backend, env in (("tcgen05", None), ("generic", "0")):
        monkeypatch.delenv("NIS_BWD_TC", raising=False)
        if env is not None:
            monkeypatch.setenv("NIS_BWD_TC", env)
        NF = make_manager(cfg)
        model = NF._model
        model.load_state_dict(sd)
        model.train()
        XJ = model(x.cuda())
        torch.var(fres * XJ[:, -1]).backward()
        grads[backend] = {k: p.grad.detach().cpu().double() for k, p in model.named_parameters()}
    gscale = max(float(v.abs().max()) for v in grads["generic"].values())
    worst = 0.0
    for k in grads["generic"]:
        close(grads["tcgen05"][k], grads["generic"][k], k, gscale, flips=True)
        worst = max(worst, float((grads["tcgen05"][k] - grads["generic"][k]).abs().max()) / gscale)
    print("tc vs generic backward at 2^%d points: worst |diff| / model gradient scale = %.2e" % (log2b, worst))


@pytest.mark.parametrize("which,B", [(3, 3000), (3, 1 << 14)], ids=["cfg5_small_3000", "cfg5_small_2p14"])
def test_activation_cache_backward_equals_recomputing_backward(monkeypatch, which, B):
    """nis_flow_forward_cached keeps z_1..z_depth of the streamed-weights layer passes and nis_flow_backward_cached reads them;
    with NIS_ACT_CACHE_MAX_BYTES=0 the backward runs the layer passes again (BatchNorm scale / shift rebuilt from the saved
    float32 batch statistics, so z differs from the forward's by a rounding).  Forward outputs are bit-identical, the
    gradients agree to 1e-5 of the model's gradient scale."""
    import ctypes
    from nf_b200 import _cabi
    cfg = dict(BIG[which], B=B)
    layers = oracle_layers(cfg)
    cells, _ = oflow.compile_layers(layers, cfg["n_flow"])
    sd = oflow.init_state_dict(cells, cfg["n_flow"], cfg["kind"], cfg["n_bins"], cfg["NN"], seed=35,
                               dtype=torch.float32, bn_jitter=0.2)
    gen = torch.Generator().manual_seed(8)
    x = torch.rand(cfg["B"], cfg["n_flow"], generator=gen, dtype=torch.float32)
    fres = torch.exp(-((x - 0.5) ** 2).sum(-1) / 0.2).cuda()
    res = {}
    for mode, env in (("cached", None), ("recompute", "0")):
        monkeypatch.delenv("NIS_ACT_CACHE_MAX_BYTES", raising=False)
        if env is not None:
            monkeypatch.setenv("NIS_ACT_CACHE_MAX_BYTES", env)
        NF = make_manager(cfg)
        model = NF._model
        model.load_state_dict(sd)
        model.train()
        spec = model.spec()
        n = spec.act_saved_count(_cabi.lib(), B)
        assert (n > 0) == (mode == "cached"), (mode, n)
        xin = x.cuda().requires_grad_(True)
        XJ = model(xin)
        torch.var(fres * XJ[:, -1]).backward()
        res[mode] = (XJ.detach().clone(), xin.grad.clone(), {k: p.grad.detach().clone() for k, p in model.named_parameters()})
    assert torch.equal(res["cached"][0], res["recompute"][0])
    gscale = max(float(g.abs().max()) for g in res["recompute"][2].values())
    xscale = float(res["recompute"][1].abs().max())
    assert gscale > 0 and xscale > 0
    # a z within a rounding of 0 may land on the other side of the ReLU in the recomputation: that point's dL/dx then
    # differs by one hidden unit's contribution -- allow a handful of such rows
    rowdiff = (res["cached"][1] - res["recompute"][1]).abs().amax(dim=1) / xscale
    bad_rows = int((rowdiff > 1e-5).sum())
    print("rows of dL/dx beyond 1e-5 of the scale: %d of %d (max %.2e)" % (bad_rows, B, float(rowdiff.max())))
    assert bad_rows <= max(2, B // 1000), bad_rows
    worst = 0.0
    for k, g in res["cached"][2].items():
        assert torch.isfinite(g).all(), k
        worst = max(worst, float((g - res["recompute"][2][k]).abs().max()) / gscale)
    print("activation cache vs recomputing backward: worst |diff| / gradient scale = %.2e" % worst)
    assert worst <= 1e-4, worst


def test_full_cfg5_backward_with_several_flushes_matches_generic(monkeypatch):
    """BASELINE configs[4] at its full shape (width 256: two column blocks per weight-gradient row block) on 2^14 points:
    16 tiles per output-layer CTA of the streamed-weights wgrad kernel = two accumulator flushes, the second a vector
    reduction onto the slice."""
    cfg = dict(name="cfg5", kind="quad", n_flow=16, n_cells=8, n_bins=64, NN=[256] * 4)
    test_tensor_core_backward_at_large_batch_matches_generic(monkeypatch, cfg, 14)


WIDE_BWD = [
    dict(name="wide_lin128", kind="lin", n_flow=8, n_pass_through=4, n_cells=4, n_bins=48, NN=[128] * 3, roll_step=4, B=2400),
    dict(name="quad64_20bins", kind="quad", n_flow=8, n_cells=6, n_bins=20, NN=[64] * 2, B=2600),
    dict(name="wide_quad128_1layer", kind="quad", n_flow=6, n_cells=6, n_bins=12, NN=[128], B=2300),
]


@pytest.mark.parametrize("cfg", WIDE_BWD, ids=[c["name"] for c in WIDE_BWD])
def test_streamed_weight_backward_shapes(cfg):
    """flow_bwd_wide.cu beyond cfg4 / cfg5: PWLin cells, width 64 with a bin count the resident-weights kernel
    does not take, a single hidden layer (z_1 stored by the head itself), ragged last tile."""
    test_gradients_match_oracle_autograd_at_size(cfg, "train")


@pytest.mark.parametrize("which", [0, 1, 3], ids=["cfg2", "cfg4", "cfg5_small"])
def test_tensor_core_backward_at_small_batches(which):
    """600 points in two minibatches of 300: tensor-core forward and backward on three tiles each."""
    test_gradients_match_oracle_autograd_at_size(dict(BIG[which], B=600), "train")
