"""Pin the CPU oracle (oracle/) against vectors dumped from the reference itself (tests/golden)."""
import math

import numpy as np
import pytest
import torch

from conftest import RAMBO_EDGE_CASES, RAMBO_PDF_CASES, FLOW_CASES, GRAD_CASES, RAMBO_CASES
from oracle import flow as oflow
from oracle import nis as onis
from oracle import rambo as orambo


def layers_for(meta):
    if meta["kind"] == "quad":
        return oflow.pwquad_layers(meta["n_flow"], meta["n_cells"])
    return oflow.pwlin_layers(meta["n_flow"], meta["n_pass_through"], meta["n_cells"], meta["roll_step"])


@pytest.mark.parametrize("case", FLOW_CASES)
def test_topology_matches_reference_children(golden, case):
    g = golden("flow_" + case)
    layers = layers_for(g.meta)
    assert [L["name"] for L in layers] == g.meta["children"]


@pytest.mark.parametrize("case", FLOW_CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_flow_forward_matches_reference(golden, case, mode):
    g = golden("flow_" + case)
    meta = g.meta
    layers = layers_for(meta)
    sd = g.state_dict()
    xj = g.t("xj")
    stats, trace = {}, {}
    XJ, bins = oflow.flow_forward(layers, sd, xj, meta["kind"], meta["n_bins"], train=(mode == "train"),
                                  stats=stats, trace=trace)
    ref = g.t(mode + "/XJ")
    assert torch.allclose(XJ, ref, rtol=1e-12, atol=1e-14), float((XJ - ref).abs().max())
    for name, v in trace.items():
        assert torch.allclose(v, g.t("%s/trace/%s" % (mode, name)), rtol=1e-12, atol=1e-14), name
    for i, b in enumerate(bins):
        if meta["kind"] != "affine":                      # (the affine cell has no bins)
            assert np.array_equal(b.numpy(), g["%s/bins/%d" % (mode, i)]), "bins cell %d" % i
    if mode == "train":
        for k in g.keys("train/stats/"):
            name = k[len("train/stats/"):]
            if name.endswith("num_batches_tracked"):
                assert int(g[k]) == int(sd[name]) + 1
            else:
                assert torch.allclose(stats[name], g.t(k), rtol=1e-12, atol=1e-15), name


@pytest.mark.parametrize("case", FLOW_CASES)
def test_compiled_index_tables_equal_layerwise(golden, case):
    g = golden("flow_" + case)
    meta = g.meta
    layers = layers_for(meta)
    cells, out_perm = oflow.compile_layers(layers, meta["n_flow"])
    sd = g.state_dict()
    xj = g.t("xj")
    for train in (False, True):
        a, ba = oflow.flow_forward(layers, sd, xj, meta["kind"], meta["n_bins"], train=train)
        b, bb = oflow.flow_forward_compiled(cells, out_perm, sd, xj, meta["kind"], meta["n_bins"], train=train)
        assert torch.equal(a, b)
        for u, v in zip(ba, bb):
            assert torch.equal(u, v)


@pytest.mark.parametrize("case", GRAD_CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_autograd_of_oracle_matches_reference_gradients(golden, case, mode):
    g = golden("flow_" + case)
    meta = g.meta
    layers = layers_for(meta)
    train = mode == "train"
    sd = g.state_dict()
    pnames = [k for k in sd if k.endswith("weight") or k.endswith("bias")]
    for k in pnames:
        sd[k].requires_grad_(True)
    xj = g.t("xj")
    # (i) variance loss, X detached (manager.py:225-258)
    XJ, _ = oflow.flow_forward(layers, sd, xj, meta["kind"], meta["n_bins"], train=train)
    fres = g.t(mode + "/grad_var/fres")
    loss = onis.minibatch_loss(fres, XJ[:, -1], fres.max(), "var")
    assert torch.allclose(loss, g.t(mode + "/grad_var/loss"), rtol=1e-11)
    grads = torch.autograd.grad(loss, [sd[k] for k in pnames], allow_unused=True)
    for k, gr in zip(pnames, grads):
        ref = g.t("%s/grad_var/%s" % (mode, k))
        gr = torch.zeros_like(ref) if gr is None else gr
        assert torch.allclose(gr, ref, rtol=1e-8, atol=1e-12 * max(1.0, float(ref.abs().max()))), k
    # (ii) generic upstream gradient
    xin = xj.clone().requires_grad_(True)
    XJ, _ = oflow.flow_forward(layers, sd, xin, meta["kind"], meta["n_bins"], train=train)
    G = g.t(mode + "/grad_lin/G")
    grads = torch.autograd.grad((XJ * G).sum(), [xin] + [sd[k] for k in pnames], allow_unused=True)
    assert torch.allclose(grads[0], g.t(mode + "/grad_lin/dxj"), rtol=1e-8, atol=1e-11)
    for k, gr in zip(pnames, grads[1:]):
        ref = g.t("%s/grad_lin/%s" % (mode, k))
        gr = torch.zeros_like(ref) if gr is None else gr
        assert torch.allclose(gr, ref, rtol=1e-8, atol=1e-11 * max(1.0, float(ref.abs().max()))), k


@pytest.mark.parametrize("case", FLOW_CASES)
@pytest.mark.parametrize("train", [False, True])
def test_inverse_flow_undoes_the_reference_pinned_forward(golden, case, train):
    """SURVEY 8 f4.  The reference has no inverse (README.md:68-69: to do), so the oracle's inverse cells are pinned through
    the forward that IS pinned to the reference: flow_inverse(flow_forward(x)) = x with Jacobian 1, on every golden model
    (all topologies: rolls, masks, extra cells), for points off the bin edges and below PWQuad's clamp at 1 - 1e-6."""
    g = golden("flow_" + case)
    m = g.meta
    layers = oflow.pwlin_layers(m["n_flow"], m["n_pass_through"], m["n_cells"], m["roll_step"]) if m["kind"] != "quad" \
        else oflow.pwquad_layers(m["n_flow"], m["n_cells"])
    sd = g.state_dict()
    x = 0.001 + 0.998 * torch.rand(300, m["n_flow"], generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    xj = torch.cat((x, torch.full((300, 1), 1.7, dtype=torch.float64)), 1)
    with torch.no_grad():
        y, bins = oflow.flow_forward(layers, sd, xj, m["kind"], m["n_bins"], train=train)
        xb, ibins = oflow.flow_inverse(layers, sd, y, m["kind"], m["n_bins"], train=train)
    assert all(torch.equal(a, b) for a, b in zip(bins, ibins))
    assert torch.allclose(xb[:, :-1], x, rtol=0, atol=1e-9)
    assert torch.allclose(xb[:, -1], xj[:, -1], rtol=1e-9)


def test_pwquad_condition_output_is_the_derivative_of_the_log_jacobian(golden):
    """The per-point sensitivity the oracle hands to the GPU parity bound (tests/gpu_util.py COND_K) is what it says:
    sum_t |d log f_t / d x_t| of one cell in eval mode, checked against autograd on a golden model."""
    g = golden("flow_quad8d_small")
    m = g.meta
    layers = oflow.pwquad_layers(m["n_flow"], m["n_cells"])
    cell = next(L for L in layers if L["type"] == "cell")
    sd = g.state_dict()
    d = m["n_flow"]
    x = (0.02 + 0.96 * torch.rand(64, d, generator=torch.Generator().manual_seed(9), dtype=torch.float64)).requires_grad_(True)
    xj = torch.cat((x, torch.ones(64, 1, dtype=torch.float64)), 1)
    cond = []
    out, _ = oflow.pwquad_cell(sd, cell["name"], xj, cell["P"], m["n_bins"], False, cond=cond)
    (grad,) = torch.autograd.grad(torch.log(out[:, -1]).sum(), x)
    assert torch.allclose(grad[:, cell["P"]:].abs().sum(-1), cond[0], rtol=1e-9)


@pytest.mark.parametrize("case", RAMBO_CASES)
def test_rambo_matches_reference(golden, case):
    g = golden("rambo_" + case)
    m = g.meta
    mom, w = orambo.generate_kinematics(m["E_cm"], g.t("r"), m["initial"], m["final"], **m["cuts"])
    ref_mom, ref_w = g.t("momenta"), g.t("weight")
    assert np.array_equal((w != 0).numpy(), (ref_w != 0).numpy()), "cut mask"
    assert torch.allclose(w, ref_w, rtol=1e-12, atol=0), float(((w - ref_w) / ref_w.abs().clamp_min(1e-300)).abs().max())
    assert torch.allclose(mom, ref_mom, rtol=1e-12, atol=1e-10 * m["E_cm"] * 1e-3)


@pytest.mark.parametrize("case", RAMBO_CASES)
def test_rambo_inverse_is_pinned_by_the_reference_momenta(golden, case):
    """SURVEY 8 f4.  The reference has no inverse phase space (README.md:68-69), so the oracle's is pinned through what the
    reference did produce: its golden momenta go back to the uniforms it was given (2e-9) and to its weight where no cut
    removed the event."""
    g = golden("rambo_" + case)
    m = g.meta
    r, w = orambo.invert_kinematics(m["E_cm"], g.t("momenta"), m["initial"], m["final"])
    assert torch.allclose(r, g.t("r"), rtol=0, atol=2e-9)
    kept = g.t("weight") != 0
    assert torch.allclose(w[kept], g.t("weight")[kept], rtol=1e-7)


@pytest.mark.parametrize("case", RAMBO_EDGE_CASES)
def test_rambo_edges_match_reference(golden, case):
    """r on the ends of [0,1]: the oracle follows the reference bit for bit, including where the reference's momenta
    overflow (compared with inf / NaN in the same places)."""
    g = golden("rambo_" + case)
    m = g.meta
    mom, w = orambo.generate_kinematics(m["E_cm"], g.t("r"), m["initial"], m["final"], **m["cuts"])
    ref_mom, ref_w = g.t("momenta"), g.t("weight")
    assert bool(torch.isfinite(ref_w).all()) and bool(torch.isfinite(w).all())
    assert np.array_equal((w != 0).numpy(), (ref_w != 0).numpy()), "cut mask"
    assert torch.allclose(w, ref_w, rtol=1e-12, atol=0)
    assert torch.allclose(mom, ref_mom, rtol=1e-12, atol=1e-13 * m["E_cm"], equal_nan=True)


@pytest.mark.parametrize("case", RAMBO_PDF_CASES)
def test_rambo_pdf_active_matches_reference(golden, case):
    """flat_phase_space_generator.py:157-187,213-219,283 with the stand-in PDF: Bjorken x sampling, PDF densities, x
    cut, per-event E_cm, lab-frame cuts, 1 / (2 x1 x2 s)."""
    from pdf_stub import StubPdf
    g = golden("rambo_" + case)
    m = g.meta
    mom, w = orambo.generate_kinematics(m["E_cm"], g.t("r"), m["initial"], m["final"], pdf=StubPdf(), pdf_active=True,
                                        tau=m["tau"], pdgs=m["pdgs"], **m["cuts"])
    ref_mom, ref_w = g.t("momenta"), g.t("weight")
    assert np.array_equal((w != 0).numpy(), (ref_w != 0).numpy()), "cut mask"
    assert torch.allclose(w, ref_w, rtol=1e-12, atol=0)
    assert torch.allclose(mom, ref_mom, rtol=1e-12, atol=1e-13 * m["E_cm"])


def test_rambo_known_answers():
    # massless 2->4 at E_cm=1000 has constant weight get_flatWeights/(2 s)
    r = torch.rand(64, 8, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    mom, w = orambo.generate_kinematics(1000.0, r, [0.0, 0.0], [0.0] * 4)
    assert torch.allclose(w, torch.full_like(w, 0.06648282151394422), rtol=1e-13)
    # momentum conservation and on-shell masses, massive case
    mom, w = orambo.generate_kinematics(1000.0, r, [100.0, 100.0], [100.0] * 4)
    assert (mom[:, :2].sum(1) - mom[:, 2:].sum(1)).abs().max() < 5e-12 * 1000
    mass = torch.sqrt(mom[:, 2:, 0] ** 2 - (mom[:, 2:, 1:] ** 2).sum(-1))
    assert (mass - 100.0).abs().max() < 1e-9
    assert torch.allclose(mom[0, 0], torch.tensor([500.0, 0, 0, math.sqrt(500.0 ** 2 - 100.0 ** 2)], dtype=torch.float64))


def test_rambo_errors():
    with pytest.raises(orambo.PhaseSpaceGeneratorError):
        orambo.generate_kinematics(1000.0, torch.full((4, 8), float("nan")), [0.0, 0.0], [0.0] * 4)
    with pytest.raises(orambo.PhaseSpaceGeneratorError):
        orambo.generate_kinematics(1000.0, torch.rand(4, 8), [0.0], [0.0] * 4)


def test_integrate_formula_and_known_answer(golden):
    g = golden("integrate_camel")
    analytic = 2 * (0.5 * math.sqrt(0.04 * math.pi) * (math.erf(3.75) + math.erf(1.25))) ** 2
    assert abs(analytic - 0.23232) < 1e-5
    # the reference's own untrained-flow estimate is consistent with the analytic value within its
    # honest error (reported error is ~sqrt(nitn) too small, manager.py:403)
    assert abs(float(g["sig_eval"]) - analytic) < 5 * math.sqrt(10) * float(g["err_eval"])
    mean = torch.tensor([1.0, 1.1, 0.9], dtype=torch.float64)
    var = torch.tensor([0.2, 0.1, 0.4], dtype=torch.float64)
    sig, err, honest = onis.integrate_combine(mean, var, 100)
    assert math.isclose(float(sig), (1 / 0.2 + 1.1 / 0.1 + 0.9 / 0.4) / (5 + 10 + 2.5))
    assert math.isclose(float(err), math.sqrt(1 / 17.5) / math.sqrt(300))
    assert math.isclose(float(honest), math.sqrt(1 / (100 * 17.5)))


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_epoch_state_machine_matches_reference_training_trace(golden, case):
    g = golden("train_" + case)
    m = g.meta
    losses = g["losses"]
    sm = onis.EpochStateMachine(float(g["int_loss"]), preburn_time=m["preburn_time"], kill_counter=m["kill_counter"])
    stopped_at = None
    for i, loss in enumerate(losses):
        if sm.step(i, float(loss)):
            stopped_at = i
            break
    # the reference logged exactly len(losses) epochs: either it ran all of them or it stopped on the last
    assert stopped_at in (None, len(losses) - 1)
    if len(losses) < m["epochs"]:
        assert stopped_at == len(losses) - 1
    assert sm.best_epoch == int(g["best_epoch"])
    assert math.isclose(float(sm.best_loss), float(g["best_loss"]), rel_tol=1e-12)
    assert math.isclose(float(g["best_func_count"]), 2 * m["batch"] * 2 + m["batch"] * len(losses))
