#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz by RUNNING THE REFERENCE (NGoetz/NF).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference.  The only accommodation is the PWLin shim
of SURVEY.md §8(c): ``PWLinManager.create_model`` raises a mixed-dtype RuntimeError on torch >= 1.8
*after* ``_model`` is built (manager.py:493-499); we catch it and call ``_model.double()``.

All weights / inputs are rounded to float32-representable values before the reference is run, so the
fp32 CUDA path and the fp64 reference see bit-identical inputs.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
from nisrep.normalizing_flows.manager import AffineManager, PWQuadManager, PWLinManager  # noqa: E402
from nisrep.PhaseSpace.flat_phase_space_generator import FlatInvertiblePhasespace  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def f32r(t):
    return t.float().double()


def camel(x):
    return torch.exp(-((x[:, 0] - 0.75) ** 2 + (x[:, 1] - 0.75) ** 2) / (0.2 ** 2)) + \
        torch.exp(-((x[:, 0] - 0.25) ** 2 + (x[:, 1] - 0.25) ** 2) / (0.2 ** 2))


def gauss(x):
    return torch.exp(-torch.sum((x - 0.5) ** 2, -1) / 0.1)


def build(kind, n_flow, seed, **kw):
    torch.manual_seed(seed)
    if kind == "quad":
        NF = PWQuadManager(n_flow=n_flow)
        NF.create_model(kw["n_cells"], kw["n_bins"], kw["NN"])
    elif kind == "affine":
        # AffineManager.create_model (manager.py:429-453) raises torch's mixed-dtype error in its trial pass, after
        # _model is assigned - the same quirk and the same shim as PWLinManager
        NF = AffineManager(n_flow=n_flow)
        try:
            NF.create_model(kw["n_pass_through"], kw["n_cells"], kw["NN"], kw["roll_step"])
        except RuntimeError:
            pass
        NF._model.double()
    else:
        NF = PWLinManager(n_flow=n_flow)
        try:
            NF.create_model(kw["n_pass_through"], kw["n_cells"], kw["n_bins"], kw["NN"], kw["roll_step"])
        except RuntimeError:
            pass
        NF._model.double()
    model = NF._model
    # make BN affine / running statistics non-trivial and everything fp32-representable
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for mod_name, mod in model.named_modules():
            if isinstance(mod, torch.nn.BatchNorm1d):
                n = mod.num_features
                mod.weight.copy_(1 + 0.3 * (torch.rand(n, generator=g, dtype=torch.float64) - 0.5))
                mod.bias.copy_(0.2 * (torch.rand(n, generator=g, dtype=torch.float64) - 0.5))
                mod.running_mean.copy_(0.2 * (torch.rand(n, generator=g, dtype=torch.float64) - 0.5))
                mod.running_var.copy_(1 + 0.4 * (torch.rand(n, generator=g, dtype=torch.float64) - 0.5))
        for p in model.parameters():
            p.copy_(f32r(p))
        for b in model.buffers():
            if b.dtype.is_floating_point:
                b.copy_(f32r(b))
    return NF


class Recorder:
    """Records the bin indices the reference computes: PWQuad via torch.argmax
    (coupling_cells.py:201), PWLin via torch.floor (coupling_cells.py:128)."""

    def __init__(self):
        self.bins = []

    def __enter__(self):
        self._argmax, self._floor = torch.argmax, torch.floor

        def argmax(*a, **k):
            r = self._argmax(*a, **k)
            self.bins.append(r.clone())
            return r

        def floor(*a, **k):
            r = self._floor(*a, **k)
            self.bins.append(r.long().clone())
            return r

        torch.argmax, torch.floor = argmax, floor
        return self

    def __exit__(self, *exc):
        torch.argmax, torch.floor = self._argmax, self._floor


def run_flow(model, xj, train):
    model.train(train)
    trace = {}
    hooks = []
    for name, mod in model.named_children():
        hooks.append(mod.register_forward_hook(
            lambda m, i, o, name=name: trace.__setitem__(name, o.detach().clone())))
    with Recorder() as rec:
        out = model(xj)
    for h in hooks:
        h.remove()
    return out, trace, rec.bins


def flow_case(name, kind, n_flow, B, seed, grad=False, **kw):
    NF = build(kind, n_flow, seed, **kw)
    model = NF._model
    g = torch.Generator().manual_seed(seed + 2)
    x = f32r(torch.rand(B, n_flow, generator=g, dtype=torch.float64))
    # edge rows: exact 0, near-1 (PWQuad clamps at 1-1e-6; PWLin has no clamp, so keep < 1)
    x[0, :] = 0.0
    x[1, :] = f32r(torch.tensor(1.0 - 1e-7)) if kind == "quad" else f32r(torch.tensor(1.0 - 1e-6))
    x[2, :] = 0.5
    jin = f32r(0.5 + torch.rand(B, 1, generator=g, dtype=torch.float64))
    xj = torch.cat((x, jin), 1)

    out = {}
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    for k, v in sd0.items():
        out["sd/" + k] = v.numpy().astype(np.int64 if not v.dtype.is_floating_point else np.float32)
    out["xj"] = xj.numpy().copy()

    # eval-mode forward (running statistics)
    with torch.no_grad():
        XJ, trace, bins = run_flow(model, xj, train=False)
    out["eval/XJ"] = XJ.numpy()
    for i, b in enumerate(bins):
        out["eval/bins/%d" % i] = b.reshape(b.shape[0], -1).numpy().astype(np.int32)
    for k, v in trace.items():
        out["eval/trace/" + k] = v.numpy()

    # train-mode forward (batch statistics; updates running stats)
    with torch.no_grad():
        XJ, trace, bins = run_flow(model, xj, train=True)
    out["train/XJ"] = XJ.numpy()
    for i, b in enumerate(bins):
        out["train/bins/%d" % i] = b.reshape(b.shape[0], -1).numpy().astype(np.int32)
    for k, v in trace.items():
        out["train/trace/" + k] = v.numpy()
    for k, v in model.state_dict().items():
        if "running" in k or "num_batches" in k:
            out["train/stats/" + k] = v.numpy().copy()

    if grad:
        model.load_state_dict(sd0)
        fun = camel if n_flow == 2 else gauss
        for mode in ("train", "eval"):
            model.train(mode == "train")
            # (i) the variance loss exactly as manager.py:225-258 builds it (X detached)
            model.zero_grad()
            XJ = model(xj)
            X = XJ[:, :-1].detach()
            fres = fun(X)
            maxf = fres.max()
            loss = torch.var(fres * XJ[:, -1] / maxf)
            loss.backward()
            out[mode + "/grad_var/loss"] = loss.detach().numpy()
            out[mode + "/grad_var/fres"] = fres.numpy()
            for k, p in model.named_parameters():
                out[mode + "/grad_var/" + k] = p.grad.numpy().copy()
            # (ii) a generic upstream gradient on every output column (exercises dL/dX too)
            model.load_state_dict(sd0)
            model.zero_grad()
            xin = xj.clone().requires_grad_(True)
            G = f32r(torch.randn(B, n_flow + 1, generator=g, dtype=torch.float64))
            XJ = model(xin)
            (XJ * G).sum().backward()
            out[mode + "/grad_lin/G"] = G.numpy()
            out[mode + "/grad_lin/dxj"] = xin.grad.numpy().copy()
            for k, p in model.named_parameters():
                out[mode + "/grad_lin/" + k] = p.grad.numpy().copy()
            model.load_state_dict(sd0)

    meta = dict(name=name, kind=kind, n_flow=n_flow, B=B, seed=seed,
                children=[n for n, _ in model.named_children()], **kw)
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "flow_%s.npz" % name), **out)
    print("flow", name, "children", meta["children"], "J mean", float(XJ[:, -1].mean()))


def rambo_case(name, initial, final, E_cm, B, seed, **cuts):
    gen = FlatInvertiblePhasespace(initial, final, pdf=None, pdf_active=False)
    g = torch.Generator().manual_seed(seed)
    r = torch.rand(B, gen.nDimPhaseSpace(), generator=g, dtype=torch.float64)
    mom, w = gen.generateKinematics_batch(E_cm, r, **cuts)
    meta = dict(name=name, initial=initial, final=final, E_cm=E_cm, B=B, seed=seed, cuts=cuts)
    np.savez_compressed(os.path.join(HERE, "rambo_%s.npz" % name), r=r.numpy(), momenta=mom.numpy(),
                        weight=w.numpy(), meta=np.array(json.dumps(meta)))
    print("rambo", name, "pass fraction", float((w != 0).double().mean()), "w[0]", float(w[0]))


def rambo_edge_case(name, initial, final, E_cm, seed, with_zero, **cuts):
    """Uniforms on the ends of [0,1]: every column in turn, and pairs of mass columns, set to
    0, denormal, 2^-24 (the float32 grid next to 0), 1-2^-24 and 1.  ``with_zero=False`` leaves 0 and the denormal
    out, so that the reference's batch-wide stopping rule (:333-348) runs its usual 120+ levels instead of 60."""
    gen = FlatInvertiblePhasespace(initial, final, pdf=None, pdf_active=False)
    n = len(final)
    nd = gen.nDimPhaseSpace()
    vals = ([0.0, 5e-324] if with_zero else []) + [2.0 ** -24, 1.0 - 2.0 ** -24, 1.0]
    rows = [(c, v, None, None) for c in range(nd) for v in vals]
    rows += [(a, va, b, vb) for a in range(n - 2) for b in range(n - 2) if a < b for va in vals for vb in vals]
    g = torch.Generator().manual_seed(seed)
    r = torch.rand(len(rows) + 8, nd, generator=g, dtype=torch.float64)
    for i, (a, va, b, vb) in enumerate(rows):
        r[i, a] = va
        if b is not None:
            r[i, b] = vb
    mom, w = gen.generateKinematics_batch(E_cm, r.clone(), **cuts)
    # the weight is finite on every row; the reference's *momenta* are not (its boost divides by sqrt(1-beta^2),
    # which rounds to 0 for a parent of mass ~2^-30 K: rows with r = 0 / denormal) - stored as they come
    assert bool(torch.isfinite(w).all()), name
    meta = dict(name=name, initial=initial, final=final, E_cm=E_cm, B=r.shape[0], seed=seed, cuts=cuts)
    np.savez_compressed(os.path.join(HERE, "rambo_%s.npz" % name), r=r.numpy(), momenta=mom.numpy(),
                        weight=w.numpy(), meta=np.array(json.dumps(meta)))
    print("rambo", name, "B", r.shape[0], "pass fraction", float((w != 0).double().mean()), "w range",
          float(w.min()), float(w.max()), "rows with non-finite momenta",
          int((~torch.isfinite(mom).all(-1).all(-1)).sum()))


def rambo_pdf_case(name, initial, final, E_cm, B, seed, tau, pdgs, **cuts):
    """pdf-active phase space (flat_phase_space_generator.py:157-187, 213-219, 283): the reference needs `import
    lhapdf` only to construct (:37-38) and then calls ``pdf.xfxQ2`` - an empty module of that name and the analytic
    stand-in of tests/pdf_stub.py let it run unmodified."""
    import types
    sys.modules.setdefault("lhapdf", types.ModuleType("lhapdf"))
    sys.path.insert(0, os.path.dirname(HERE))
    from pdf_stub import StubPdf
    gen = FlatInvertiblePhasespace(initial, final, pdf=StubPdf(), pdf_active=True, tau=tau)
    g = torch.Generator().manual_seed(seed)
    r = torch.rand(B, gen.nDimPhaseSpace() + 2, generator=g, dtype=torch.float64)
    if not tau:
        # the last two columns ARE x2, x1: keep sqrt(x1 x2) E_cm above the final-state masses
        lo = (sum(final) / E_cm) ** 2 + 0.02
        r[:, -2:] = lo ** 0.5 + (1.0 - lo ** 0.5) * r[:, -2:]
    mom, w = gen.generateKinematics_batch(E_cm, r.clone(), pdgs=list(pdgs), **cuts)
    assert bool(torch.isfinite(w).all())
    meta = dict(name=name, initial=initial, final=final, E_cm=E_cm, B=B, seed=seed, cuts=cuts, tau=tau, pdgs=list(pdgs))
    np.savez_compressed(os.path.join(HERE, "rambo_%s.npz" % name), r=r.numpy(), momenta=mom.numpy(),
                        weight=w.numpy(), meta=np.array(json.dumps(meta)))
    print("rambo", name, "pass fraction", float((w != 0).double().mean()), "w range", float(w.min()), float(w.max()))


class FakeRun:
    """Stand-in for the Sacred run object the reference logs to (manager.py:89-93,197-198,286-288)."""
    _id = "run"

    def __init__(self):
        import datetime
        self.start_time = datetime.datetime.utcnow()
        self.scalars = {}

    def log_scalar(self, name, val, step):
        self.scalars.setdefault(name, []).append((int(step), float(val)))


def train_case(name, seed, epochs, batch, mini, **kw):
    NF = build("quad", 2, seed, n_cells=2, n_bins=4, NN=[3] * 3)
    NF._model.train()
    optim = torch.optim.Adamax(NF._model.parameters(), lr=2e-2, weight_decay=1e-4)
    run = FakeRun()
    torch.manual_seed(seed + 7)
    with tempfile.TemporaryDirectory() as td:
        ret = NF._train_variance_forward_seq(camel, optim, True, td, batch, epochs, 0, False, True,
                                             run=run, mini_batch_size=mini, **kw)
    losses = [v for _, v in run.scalars["training.loss"]]
    out = dict(losses=np.array(losses), int_loss=np.array(float(NF.int_loss)),
               best_loss=np.array(float(NF.best_loss)), best_epoch=np.array(int(NF.best_epoch)),
               best_func_count=np.array(float(NF.best_func_count)),
               best_loss_rel=np.array(float(NF.best_loss_rel)),
               meta=np.array(json.dumps(dict(name=name, seed=seed, epochs=epochs, batch=batch, mini=mini, **kw))))
    np.savez_compressed(os.path.join(HERE, "train_%s.npz" % name), **out)
    print("train", name, "epochs run", len(losses), "best_epoch", int(NF.best_epoch), "ret", ret)


def integrate_case(seed):
    NF = build("quad", 2, seed, n_cells=2, n_bins=4, NN=[3] * 3)
    torch.manual_seed(seed + 3)
    sig, err = NF.integrate(camel, 10, 10000)
    NF._model.eval()
    torch.manual_seed(seed + 3)
    sig_e, err_e = NF.integrate(camel, 10, 10000)
    sd = {"sd/" + k: v.numpy().astype(np.int64 if not v.dtype.is_floating_point else np.float32)
          for k, v in NF._model.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "integrate_camel.npz"), sig=np.array(float(sig)),
                        err=np.array(float(err)), sig_eval=np.array(float(sig_e)), err_eval=np.array(float(err_e)),
                        analytic=np.array(0.232322), **sd)
    print("integrate camel (untrained, train-mode BN)", float(sig), float(err), "eval", float(sig_e), float(err_e))


if __name__ == "__main__":
    what = sys.argv[1:] or ["flow", "rambo", "train", "integrate"]
    if "flow" in what:
        flow_case("quad2d", "quad", 2, 256, 11, grad=True, n_cells=2, n_bins=4, NN=[3] * 3)
        flow_case("quad3d", "quad", 3, 192, 12, grad=True, n_cells=1, n_bins=5, NN=[8, 8])
        flow_case("quad7d", "quad", 7, 128, 13, n_cells=2, n_bins=6, NN=[16] * 2)
        flow_case("quad8d", "quad", 8, 128, 14, n_cells=6, n_bins=32, NN=[64] * 3)
        flow_case("quad8d_small", "quad", 8, 160, 15, grad=True, n_cells=2, n_bins=8, NN=[16] * 2)
        flow_case("quad9d_extra", "quad", 9, 64, 16, n_cells=10, n_bins=4, NN=[8])
        flow_case("quad16d", "quad", 16, 96, 17, n_cells=8, n_bins=8, NN=[32] * 2)
        flow_case("lin8d", "lin", 8, 128, 18, n_pass_through=4, n_cells=6, n_bins=32, NN=[64] * 3, roll_step=4)
        flow_case("lin4d", "lin", 4, 200, 19, grad=True, n_pass_through=2, n_cells=3, n_bins=10, NN=[8, 8], roll_step=1)
        flow_case("lin5d", "lin", 5, 96, 20, n_pass_through=1, n_cells=4, n_bins=7, NN=[12], roll_step=2)
    cuts = dict(pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5)
    if "affine" in what or "flow" in what:
        flow_case("affine4d", "affine", 4, 200, 21, grad=True, n_pass_through=2, n_cells=3, NN=[8, 8], roll_step=1)
        flow_case("affine6d", "affine", 6, 128, 22, n_pass_through=3, n_cells=4, NN=[16], roll_step=2)
    if "rambo" in what:
        rambo_case("m4_cuts", [100.0] * 2, [100.0] * 4, 1000.0, 512, 31, **cuts)
        rambo_case("m4_nocuts", [100.0] * 2, [100.0] * 4, 1000.0, 256, 32)
        rambo_case("m0_4", [0.0] * 2, [0.0] * 4, 1000.0, 256, 33, **cuts)
        rambo_case("m2", [100.0] * 2, [100.0] * 2, 1000.0, 128, 34, pT_mincut=50, delR_mincut=0, rap_maxcut=1.5)
        rambo_case("mixed3", [0.0, 0.0], [0.0, 50.0, 173.0], 500.0, 256, 35, pT_mincut=10, delR_mincut=0.4, rap_maxcut=-1)
        rambo_case("m5_cuts", [50.0, 100.0], [10.0, 20.0, 30.0, 40.0, 50.0], 2000.0, 256, 36, **cuts)
        rambo_case("m0_6", [0.0] * 2, [0.0] * 6, 13000.0, 128, 37, pT_mincut=30, delR_mincut=0.4, rap_maxcut=4.0)
        rambo_case("readme", [100.0] * 2, [100.0] * 4, 1000.0, 128, 38, pT_mincut=0, delR_mincut=0, rap_maxcut=-1)
    if "rambo_edges" in what or "rambo" in what:
        for z in (True, False):
            t = "edge0" if z else "edge1"
            rambo_edge_case(t + "_m0_4", [0.0] * 2, [0.0] * 4, 1000.0, 61, z)
            rambo_edge_case(t + "_m4", [100.0] * 2, [100.0] * 4, 1000.0, 62, z)
            rambo_edge_case(t + "_m4_cuts", [100.0] * 2, [100.0] * 4, 1000.0, 63, z, **cuts)
            rambo_edge_case(t + "_m0_4_cuts", [0.0] * 2, [0.0] * 4, 1000.0, 64, z, **cuts)
            rambo_edge_case(t + "_m5", [50.0, 100.0], [10.0, 20.0, 30.0, 40.0, 50.0], 2000.0, 65, z)
            rambo_edge_case(t + "_m0_3", [0.0] * 2, [0.0, 0.0, 0.0], 500.0, 66, z)
    if "rambo_pdf" in what or "rambo" in what:
        rambo_pdf_case("pdf_tau_m0_3", [0.0] * 2, [0.0] * 3, 13000.0, 256, 71, True, (21, 2), **cuts)
        rambo_pdf_case("pdf_tau_m4", [0.938, 0.938], [10.0, 20.0, 30.0, 40.0], 13000.0, 256, 72, True, (2, -1), **cuts)
        rambo_pdf_case("pdf_x_m0_4", [0.0] * 2, [0.0] * 4, 13000.0, 256, 73, False, (21, 21), **cuts)
        rambo_pdf_case("pdf_x_m3_nocuts", [0.0] * 2, [0.0, 80.4, 91.2], 1000.0, 192, 74, False, (1, -1))
        rambo_pdf_case("pdf_tau_nopdf", [0.0] * 2, [0.0] * 4, 13000.0, 128, 75, True, (0, 22), pT_mincut=20, delR_mincut=-1, rap_maxcut=2.5)
    if "train" in what:
        train_case("a", 41, 40, 2000, 1000, preburn_time=10, kill_counter=7)
        train_case("b", 42, 60, 2000, 1000, preburn_time=5, kill_counter=1)
        train_case("c", 43, 30, 1000, 1000, preburn_time=0, kill_counter=2)
    if "integrate" in what:
        integrate_case(51)
