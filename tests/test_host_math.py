"""CPU check of the product's scalar device code.

nf_b200/csrc/spline.cuh and rambo_core.cuh are plain scalar C++; tests/host/host_math.cpp builds the
very same source with g++ and this file compares it with the oracle (forward values, bin indices,
hand-derived backward vs torch.autograd of the oracle, RAMBO events vs the golden vectors).  This is a
test harness: the product never runs this host build.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import RAMBO_CASES, RAMBO_EDGE_CASES, RAMBO_PDF_CASES, ROOT, rambo_edge_rows, rambo_edge_weight_rtol
from oracle import rambo as orambo

SRC = os.path.join(ROOT, "tests", "host", "host_math.cpp")
F = ctypes.POINTER(ctypes.c_float)
D = ctypes.POINTER(ctypes.c_double)
I = ctypes.POINTER(ctypes.c_int)


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("host") / "host_math.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", so, SRC])
    return ctypes.CDLL(so)


def fp(a, t=F):
    return a.ctypes.data_as(t)


def oracle_pwlin(z, x, nb):
    Q = torch.exp(z)
    Qs = torch.cumsum(Q, -1)
    norm = Qs[:, -1:]
    Qn = Q / (norm / nb)
    C = torch.cat((torch.zeros_like(norm), Qs / norm), -1)
    a = x * nb
    k = torch.floor(a).long().clamp(0, nb - 1)
    alpha = (a - k) / nb
    Qk = torch.gather(Qn, -1, k.unsqueeze(-1)).squeeze(-1)
    y = Qk * alpha + torch.gather(C, -1, k.unsqueeze(-1)).squeeze(-1)
    return y, Qk, k


def oracle_pwquad(z, x, nb):
    from oracle.flow import pwquad_cell  # noqa: F401  (documented source of the formulas below)
    xb = torch.where(x > 1 - 1e-6, torch.full_like(x, 1 - 1e-6), x)
    V = torch.exp(z[:, :nb + 1])
    W = torch.exp(z[:, nb + 1:])
    Ws = torch.cumsum(W, -1)
    Wn = Ws[:, -1:]
    W = W / Wn
    Ws = Ws / Wn
    area = torch.cumsum((V[:, :-1] + V[:, 1:]) / 2 * W, -1)
    V = V / area[:, -1:]
    E = torch.cat((torch.zeros_like(Wn), Ws), -1)
    k = (Ws <= xb.unsqueeze(-1)).sum(-1, keepdim=True)
    Wk = torch.gather(W, -1, k).squeeze(-1)
    alpha = (xb - torch.gather(E, -1, k).squeeze(-1)) / Wk
    S = torch.cat((torch.zeros_like(Wn), torch.cumsum((V[:, :-1] + V[:, 1:]) / 2 * W, -1)), -1)
    Vk = torch.gather(V, -1, k).squeeze(-1)
    Vk1 = torch.gather(V, -1, k + 1).squeeze(-1)
    y = alpha ** 2 / 2 * ((Vk1 - Vk) * Wk) + alpha * Vk * Wk + torch.gather(S, -1, k).squeeze(-1)
    f = torch.lerp(Vk, Vk1, alpha)
    return y, f, k.squeeze(-1)


@pytest.mark.parametrize("nb", [1, 4, 7, 32])
def test_pwlin_forward_backward(host, nb):
    g = torch.Generator().manual_seed(nb)
    n = 400
    z = (2.0 * torch.randn(n, nb, generator=g)).float()
    x = torch.rand(n, generator=g).float()
    x[0] = 0.0
    gy = torch.randn(n, generator=g).float()
    gJJ = torch.randn(n, generator=g).float()
    zd = z.double().requires_grad_(True)
    xd = x.double().requires_grad_(True)
    y, f, k = oracle_pwlin(zd, xd, nb)
    # L = sum gy*y + gJJ*log f   (d/df of gJJ*log f = gJJ/f, i.e. gJJ = dL/df * f)
    (gy.double() * y + gJJ.double() * torch.log(f)).sum().backward()
    zz = np.ascontiguousarray(z.numpy().copy())
    yo, fo, dx = (np.zeros(n, np.float32) for _ in range(3))
    ko = np.zeros(n, np.int32)
    host.host_pwlin(n, nb, fp(zz), fp(x.numpy()), fp(gy.numpy()), fp(gJJ.numpy()), fp(yo), fp(fo), fp(ko, I), fp(dx))
    assert np.array_equal(ko, k.numpy())
    np.testing.assert_allclose(yo, y.detach().numpy(), rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(fo, f.detach().numpy(), rtol=2e-6)
    np.testing.assert_allclose(zz, zd.grad.numpy(), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(dx, xd.grad.numpy(), rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("nb", [1, 2, 5, 32, 64])
def test_pwquad_forward_backward(host, nb):
    g = torch.Generator().manual_seed(100 + nb)
    n = 400
    K = 2 * nb + 1
    z = (1.5 * torch.randn(n, K, generator=g)).float()
    x = torch.rand(n, generator=g).float()
    x[0] = 0.0
    x[1] = 1.0
    x[2] = float(np.float32(1 - 1e-7))
    gy = torch.randn(n, generator=g).float()
    gf = torch.randn(n, generator=g).float()
    zd = z.double().requires_grad_(True)
    xd = x.double().requires_grad_(True)
    y, f, k = oracle_pwquad(zd, xd, nb)
    (gy.double() * y + gf.double() * f).sum().backward()
    zz = np.ascontiguousarray(z.numpy().copy())
    yo, fo, dx = (np.zeros(n, np.float32) for _ in range(3))
    ko = np.zeros(n, np.int32)
    host.host_pwquad(n, nb, fp(zz), fp(x.numpy()), fp(gy.numpy()), fp(gf.numpy()), fp(yo), fp(fo), fp(ko, I), fp(dx))
    # a bin may legitimately differ only when x sits within fp32 rounding of an edge
    bad = np.nonzero(ko != k.numpy())[0]
    assert len(bad) <= 1, bad
    good = np.setdiff1d(np.arange(n), bad)
    np.testing.assert_allclose(yo[good], y.detach().numpy()[good], rtol=5e-6, atol=5e-7)
    np.testing.assert_allclose(fo[good], f.detach().numpy()[good], rtol=1e-5)
    gref = zd.grad.numpy()
    scale = np.abs(gref).max(1, keepdims=True) + 1e-6
    assert np.max(np.abs(zz[good] - gref[good]) / scale[good]) < 2e-4
    np.testing.assert_allclose(dx[good], xd.grad.numpy()[good], rtol=5e-4, atol=1e-5)


@pytest.mark.parametrize("e", [1, 2, 3, 4, 5, 6])
def test_rambo_root_solves_the_mass_polynomial(host, e):
    host.host_rambo_root.restype = ctypes.c_double
    host.host_rambo_root.argtypes = [ctypes.c_int, ctypes.c_double]
    rs = np.concatenate([np.random.default_rng(e).random(2000), [1e-300, 1e-14, 1e-9, 1e-4, 0.5, 1 - 1e-9, 1 - 1e-15]])
    for r in rs:
        u = host.host_rambo_root(e, float(r))
        assert 0.0 <= u <= 1.0
        back = (e + 1) * u ** e - e * u ** (e + 1)
        assert abs(back - r) <= 6e-15 * max(r, 1e-300) + 2e-16, (e, r, u, back)
    # agrees with the reference's lattice bisection (oracle.bisect) to its absolute resolution
    n = e + 2
    v = torch.rand(64, n - 2, generator=torch.Generator().manual_seed(e), dtype=torch.float64)
    ub = orambo.bisect(v, n)
    ours = np.array([host.host_rambo_root(e, float(x)) for x in v[:, 0]])
    np.testing.assert_allclose(ours, ub[:, 0].numpy(), rtol=0, atol=5e-15)  # bisection is limited by the rounding of its own polynomial evaluation


class RamboDesc(ctypes.Structure):
    _fields_ = [("n_final", ctypes.c_int32), ("initial_masses", ctypes.c_double * 2),
                ("final_masses", ctypes.c_double * 8), ("E_cm", ctypes.c_double), ("pT_mincut", ctypes.c_double),
                ("delR_mincut", ctypes.c_double), ("rap_maxcut", ctypes.c_double),
                ("pdf_active", ctypes.c_int32), ("tau_mode", ctypes.c_int32),
                ("tau_min", ctypes.c_double), ("x_cut", ctypes.c_double),
                ("pdf_grid", ctypes.c_void_p * 2), ("pdf_nodes", ctypes.c_int32), ("pdf_lnx_lo", ctypes.c_double)]


@pytest.mark.parametrize("case", RAMBO_CASES)
def test_rambo_event_matches_reference_golden(host, golden, case):
    g = golden("rambo_" + case)
    m = g.meta
    n = len(m["final"])
    cuts = dict(pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1)
    cuts.update(m["cuts"])
    d = RamboDesc(n, (ctypes.c_double * 2)(*m["initial"]), (ctypes.c_double * 8)(*(m["final"] + [0.0] * (8 - n))),
                  m["E_cm"], cuts["pT_mincut"], cuts["delR_mincut"], cuts["rap_maxcut"])
    r = np.ascontiguousarray(g["r"])
    B = r.shape[0]
    mom = np.zeros((B, n + 2, 4))
    w = np.zeros(B)
    ok = np.zeros(B, np.uint8)
    assert host.host_rambo(ctypes.byref(d), ctypes.c_longlong(B), fp(r, D), fp(mom, D), fp(w, D),
                           ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))) == 0
    ref_w, ref_mom = g["weight"], g["momenta"]
    assert np.array_equal(ok.astype(bool), ref_w != 0), "cut mask"
    np.testing.assert_allclose(w, ref_w, rtol=1e-9)
    np.testing.assert_allclose(mom, ref_mom, rtol=1e-9, atol=1e-9 * m["E_cm"])


@pytest.mark.parametrize("case", RAMBO_CASES)
def test_rambo_inverse_recovers_the_reference_uniforms(host, golden, case):
    """SURVEY 8 f4 (the reference has no inverse, README.md:68-69): the kernel source's inverse map, fed the momenta the
    REFERENCE produced, returns the uniforms the reference was given and - where no cut removed the event - its weight."""
    g = golden("rambo_" + case)
    m = g.meta
    n = len(m["final"])
    d = RamboDesc(n, (ctypes.c_double * 2)(*m["initial"]), (ctypes.c_double * 8)(*(m["final"] + [0.0] * (8 - n))),
                  m["E_cm"], -1.0, -1.0, -1.0)
    mom = np.ascontiguousarray(g["momenta"])
    B = mom.shape[0]
    r = np.zeros((B, 3 * n - 4))
    w = np.zeros(B)
    assert host.host_rambo_invert(ctypes.byref(d), ctypes.c_longlong(B), fp(mom, D), fp(r, D), fp(w, D)) == 0
    np.testing.assert_allclose(r, g["r"], rtol=0, atol=2e-9)
    kept = g["weight"] != 0
    np.testing.assert_allclose(w[kept], g["weight"][kept], rtol=1e-7)
    r2, w2 = orambo.invert_kinematics(m["E_cm"], torch.as_tensor(mom), m["initial"], m["final"])
    np.testing.assert_allclose(r, r2.numpy(), rtol=0, atol=1e-10)
    np.testing.assert_allclose(w, w2.numpy(), rtol=1e-8)


def _host_rambo_case(host, g):
    m = g.meta
    n = len(m["final"])
    cuts = dict(pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1)
    cuts.update(m["cuts"])
    d = RamboDesc(n, (ctypes.c_double * 2)(*m["initial"]), (ctypes.c_double * 8)(*(m["final"] + [0.0] * (8 - n))),
                  m["E_cm"], cuts["pT_mincut"], cuts["delR_mincut"], cuts["rap_maxcut"])
    r = np.ascontiguousarray(g["r"])
    B = r.shape[0]
    mom = np.zeros((B, n + 2, 4))
    w = np.zeros(B)
    ok = np.zeros(B, np.uint8)
    assert host.host_rambo(ctypes.byref(d), ctypes.c_longlong(B), fp(r, D), fp(mom, D), fp(w, D),
                           ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))) == 0
    return mom, w, ok


@pytest.mark.parametrize("case", RAMBO_EDGE_CASES)
def test_rambo_event_on_the_ends_of_the_unit_interval(host, golden, case):
    """mass-dimension uniforms of exactly 0 / 1 (float32 uniforms hit 0 with p = 2^-24) used to
    give 0/0.  The reference's weight is finite there (u = 2^-60 resp. 1 - 2^-27/e); ours must be the same number."""
    g = golden("rambo_" + case)
    mom, w, ok = _host_rambo_case(host, g)
    ref_w, ref_mom = g["weight"], g["momenta"]
    assert np.isfinite(w).all() and np.isfinite(mom).all()
    rows = rambo_edge_rows(ref_mom, g["r"], len(g.meta["final"]))
    assert np.array_equal(ok.astype(bool)[rows], (ref_w != 0)[rows]), "cut mask"
    # rows on which the reference's momenta overflowed: its cuts compared NaNs (all pass); only the pre-cut weight
    # is comparable there, and only when we did not cut the event
    cmp = rows | (ok.astype(bool) & (ref_w != 0))
    rt = rambo_edge_weight_rtol(g["r"], g.meta)
    bad = cmp & ~(np.abs(w - ref_w) <= rt * np.abs(ref_w))
    assert not bad.any(), (np.nonzero(bad)[0], (w / ref_w - 1)[bad], rt[bad])
    assert np.isfinite(rt[cmp]).mean() > 0.8
    np.testing.assert_allclose(mom[rows], ref_mom[rows], rtol=1e-9, atol=1e-9 * g.meta["E_cm"])
    assert cmp.sum() >= 0.5 * len(w)


@pytest.mark.parametrize("case", RAMBO_PDF_CASES)
def test_rambo_event_pdf_active_matches_reference(host, golden, case):
    """The pdf-active path of the kernel source (host build) against vectors dumped from the reference with the stand-in
    PDF: Bjorken-x sampling, densities from the interpolation grid, per-event energy, lab-frame cuts."""
    from pdf_stub import StubPdf
    from nf_b200.PhaseSpace.pdf_grid import PdfGrid, X_CUT, is_parton
    g = golden("rambo_" + case)
    m = g.meta
    n = len(m["final"])
    cuts = dict(pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1)
    cuts.update(m["cuts"])
    d = RamboDesc(n, (ctypes.c_double * 2)(*m["initial"]), (ctypes.c_double * 8)(*(m["final"] + [0.0] * (8 - n))),
                  m["E_cm"], cuts["pT_mincut"], cuts["delR_mincut"], cuts["rap_maxcut"])
    d.pdf_active, d.tau_mode = 1, int(m["tau"])
    d.tau_min, d.x_cut = (max(sum(m["final"]), 1.0) / m["E_cm"]) ** 2, X_CUT
    grids = []
    for i, pdg in enumerate(m["pdgs"]):
        if is_parton(pdg):
            gr = PdfGrid(StubPdf(), pdg)
            grids.append(gr.host.numpy())
            d.pdf_grid[i] = grids[-1].ctypes.data
            d.pdf_nodes, d.pdf_lnx_lo = gr.n_nodes, gr.lnx_lo
    r = np.ascontiguousarray(g["r"])
    B = r.shape[0]
    mom = np.zeros((B, n + 2, 4))
    w = np.zeros(B)
    ok = np.zeros(B, np.uint8)
    assert host.host_rambo(ctypes.byref(d), ctypes.c_longlong(B), fp(r, D), fp(mom, D), fp(w, D),
                           ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))) == 0
    ref_w, ref_mom = g["weight"], g["momenta"]
    assert np.array_equal(ok.astype(bool), ref_w != 0), "cut mask"
    np.testing.assert_allclose(mom, ref_mom, rtol=1e-9, atol=1e-9 * m["E_cm"])
    # the densities are interpolated from 16384 nodes in ln x: 1e-8 away from x = 1, where x f ~ (1-x)^b vanishes
    np.testing.assert_allclose(w, ref_w, rtol=1e-7, atol=1e-12 * np.abs(ref_w).max())


@pytest.mark.parametrize("kind,nb", [(0, 4), (0, 32), (1, 4), (1, 32), (1, 64)])
def test_inverse_splines_round_trip_and_match_the_oracle(host, kind, nb):
    """spline.cuh's pwlin_inv / pwquad_inv (the reference has no inverse: README.md:68-69 lists it as to do)
    undo the forward maps of the same source - same bin, x back to float32 accuracy, density at x equal to the forward's -
    and agree with the float64 inverse cells of oracle/flow.py."""
    g = torch.Generator().manual_seed(7 * nb + kind)
    n = 500
    K = nb if kind == 0 else 2 * nb + 1
    z = (1.5 * torch.randn(n, K, generator=g)).float()
    x = (0.001 + 0.998 * torch.rand(n, generator=g)).float()
    yo, fo, dx = (np.zeros(n, np.float32) for _ in range(3))
    ko = np.zeros(n, np.int32)
    zz = np.ascontiguousarray(z.numpy().copy())
    zero = np.zeros(n, np.float32)
    (host.host_pwlin if kind == 0 else host.host_pwquad)(n, nb, fp(zz), fp(x.numpy()), fp(zero), fp(zero), fp(yo), fp(fo),
                                                          fp(ko, I), fp(dx))
    xi, fi = np.zeros(n, np.float32), np.zeros(n, np.float32)
    ki = np.zeros(n, np.int32)
    host.host_spline_inv(kind, n, nb, fp(np.ascontiguousarray(z.numpy())), fp(yo), fp(xi), fp(fi), fp(ki, I))
    same = ki == ko                                          # a y within float32 rounding of an edge may land next door
    assert same.mean() > 0.99
    # y is a float32: x comes back to delta_y / f (a low-density bin stretches the rounding of y)
    assert (np.abs(xi - x.numpy())[same] * fo[same]).max() < 1e-6
    np.testing.assert_allclose(fi[same], fo[same], rtol=2e-3)
    # float64 oracle inverse on the same logits
    from oracle import flow as oflow
    zd, yd = z.double(), torch.from_numpy(yo).double()
    if kind == 0:
        Q = torch.exp(zd)
        Qs = torch.cumsum(Q, -1)
        norm = Qs[:, -1:]
        Qn = Q / (norm / nb)
        C = torch.cat((torch.zeros_like(norm), Qs / norm), -1)
        k = (C[:, 1:-1] <= yd.unsqueeze(-1)).sum(-1).clamp(0, nb - 1)
        Qk = torch.gather(Qn, -1, k.unsqueeze(-1)).squeeze(-1)
        xo = k.double() / nb + (yd - torch.gather(C, -1, k.unsqueeze(-1)).squeeze(-1)) / Qk
        fo64 = Qk
    else:
        yq, fq, kq = oracle_pwquad(zd, x.double(), nb)       # forward oracle at the true x: inverse must return (x, f)
        xo, fo64, k = x.double(), fq, kq
    ok = same & (ki == k.numpy())
    assert ok.mean() > 0.98
    assert (np.abs(xi - xo.numpy())[ok] * fo64.numpy()[ok]).max() < 2e-6
    np.testing.assert_allclose(fi[ok], fo64.numpy()[ok], rtol=2e-3)
