"""GPU parity of the fused RAMBO kernel (FlatInvertiblePhasespace -> nis_rambo_generate)."""
import math

import numpy as np
import pytest
import torch

from conftest import RAMBO_CASES, RAMBO_EDGE_CASES, RAMBO_PDF_CASES, rambo_edge_rows, rambo_edge_weight_rtol
from oracle import rambo as orambo

from nf_b200.PhaseSpace.flat_phase_space_generator import FlatInvertiblePhasespace

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", RAMBO_CASES)
@pytest.mark.parametrize("where", ["cuda", "cpu"])
def test_matches_reference_golden(golden, case, where):
    g = golden("rambo_" + case)
    m = g.meta
    ps = FlatInvertiblePhasespace(m["initial"], m["final"], pdf=None, pdf_active=False)
    r = g.t("r").to(where)
    mom, w, mask = ps.generateKinematics_batch(m["E_cm"], r, return_cutmask=True, **m["cuts"])
    assert mom.device.type == where and w.dtype == torch.float64
    ref_w, ref_mom = g.t("weight"), g.t("momenta")
    assert np.array_equal(mask.cpu().numpy().astype(bool), (ref_w != 0).numpy()), "cut mask must be bit-exact"
    assert np.array_equal((w.cpu() != 0).numpy(), (ref_w != 0).numpy())
    assert torch.allclose(w.cpu(), ref_w, rtol=1e-9, atol=0)
    assert torch.allclose(mom.cpu(), ref_mom, rtol=1e-9, atol=1e-9 * m["E_cm"])


@pytest.mark.parametrize("case", RAMBO_EDGE_CASES)
def test_ends_of_the_unit_interval_match_reference(golden, case):
    """Uniforms of exactly 0 / denormal / 2^-24 / 1-2^-24 / 1 in every column (reference-dumped fixture): weights
    finite and equal to the reference's; momenta and cut masks equal wherever the reference's own momenta are finite
    (its boost overflows for a parent of mass ~2^-30 K; see conftest.rambo_edge_rows)."""
    g = golden("rambo_" + case)
    m = g.meta
    ps = FlatInvertiblePhasespace(m["initial"], m["final"], pdf=None, pdf_active=False)
    mom, w, mask = ps.generateKinematics_batch(m["E_cm"], g.t("r").cuda(), return_cutmask=True, **m["cuts"])
    mom, w, mask = mom.cpu().numpy(), w.cpu().numpy(), mask.cpu().numpy().astype(bool)
    ref_w, ref_mom = g["weight"], g["momenta"]
    assert np.isfinite(w).all() and np.isfinite(mom).all()
    rows = rambo_edge_rows(ref_mom, g["r"], len(g.meta["final"]))
    assert np.array_equal(mask[rows], (ref_w != 0)[rows]), "cut mask must be bit-exact"
    cmp = rows | (mask & (ref_w != 0))
    rt = rambo_edge_weight_rtol(g["r"], g.meta)
    bad = cmp & ~(np.abs(w - ref_w) <= rt * np.abs(ref_w))
    assert not bad.any(), (np.nonzero(bad)[0], (w / ref_w - 1)[bad], rt[bad])
    assert np.isfinite(rt[cmp]).mean() > 0.8
    np.testing.assert_allclose(mom[rows], ref_mom[rows], rtol=1e-9, atol=1e-9 * m["E_cm"])
    # the float32 grid: the same uniforms as float32 input give the same events
    mom32, w32 = ps.generateKinematics_batch(m["E_cm"], g.t("r").float().cuda(), **m["cuts"])
    r32 = g.t("r").float().double()
    same = (r32 == g.t("r")).all(1).numpy()
    assert np.isfinite(w32.cpu().numpy()).all()
    assert np.array_equal(w32.cpu().numpy()[same], w[same])


@pytest.mark.parametrize("case", RAMBO_PDF_CASES)
def test_pdf_active_matches_reference(golden, case):
    """pdf-active phase space through the public API (same constructor arguments as the reference, any object with
    ``xfxQ2``): vectors dumped from the reference with the stand-in PDF of tests/pdf_stub.py.  Cut masks bit-exact
    (the cuts act on lab-frame momenta), CM-frame momenta to 1e-9, weights to 1e-7 (densities are interpolated)."""
    from pdf_stub import StubPdf
    g = golden("rambo_" + case)
    m = g.meta
    ps = FlatInvertiblePhasespace(m["initial"], m["final"], pdf=StubPdf(), pdf_active=True, tau=m["tau"])
    assert ps.nDimInput() == g["r"].shape[1]
    mom, w, mask = ps.generateKinematics_batch(m["E_cm"], g.t("r").cuda(), pdgs=m["pdgs"], return_cutmask=True, **m["cuts"])
    ref_w, ref_mom = g.t("weight"), g.t("momenta")
    assert np.array_equal(mask.cpu().numpy().astype(bool), (ref_w != 0).numpy()), "cut mask must be bit-exact"
    assert torch.allclose(mom.cpu(), ref_mom, rtol=1e-9, atol=1e-9 * m["E_cm"])
    assert torch.allclose(w.cpu(), ref_w, rtol=1e-7, atol=1e-12 * float(ref_w.abs().max()))
    # weight-only mode and float32 uniforms run the same events
    _, w2 = ps.generateKinematics_batch(m["E_cm"], g.t("r").cuda(), pdgs=m["pdgs"], momenta=False, **m["cuts"])
    assert torch.equal(w2, w)


def test_pdf_active_at_size_matches_oracle():
    from pdf_stub import StubPdf
    B = 1 << 15
    r = torch.rand(B, 10, generator=torch.Generator().manual_seed(8), dtype=torch.float64)
    cuts = dict(pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5)
    ps = FlatInvertiblePhasespace([0.0] * 2, [0.0, 0.0, 80.4, 91.2], pdf=StubPdf(), pdf_active=True, tau=True)
    mom, w, mask = ps.generateKinematics_batch(13000.0, r.cuda(), pdgs=[21, 2], return_cutmask=True, **cuts)
    rmom, rw = orambo.generate_kinematics(13000.0, r, [0.0] * 2, [0.0, 0.0, 80.4, 91.2], pdf=StubPdf(), pdf_active=True,
                                          tau=True, pdgs=[21, 2], **cuts)
    assert np.array_equal(mask.cpu().numpy().astype(bool), (rw != 0).numpy())
    assert torch.allclose(mom.cpu(), rmom, rtol=1e-8, atol=1e-6)
    assert torch.allclose(w.cpu(), rw, rtol=1e-7, atol=1e-12 * float(rw.abs().max()))


def test_float32_uniforms_and_weight_only_mode():
    ps = FlatInvertiblePhasespace([100.0] * 2, [100.0] * 4)
    r32 = torch.rand(4096, 8, device="cuda", dtype=torch.float32)
    mom, w = ps.generateKinematics_batch(1000.0, r32, pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5)
    mom2, w2 = ps.generateKinematics_batch(1000.0, r32.double(), pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5)
    assert torch.equal(w, w2) and torch.equal(mom, mom2)
    none, w3 = ps.generateKinematics_batch(1000.0, r32, pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5, momenta=False)
    assert none is None and torch.equal(w3, w)


@pytest.mark.parametrize("n,masses", [(4, 100.0), (4, 0.0), (3, 50.0), (6, 10.0)])
def test_matches_oracle_at_size(n, masses):
    B = 1 << 15
    r = torch.rand(B, 3 * n - 4, generator=torch.Generator().manual_seed(n), dtype=torch.float64)
    cuts = dict(pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5)
    ps = FlatInvertiblePhasespace([masses] * 2, [masses] * n)
    mom, w, mask = ps.generateKinematics_batch(1000.0, r.cuda(), return_cutmask=True, **cuts)
    rmom, rw = orambo.generate_kinematics(1000.0, r, [masses] * 2, [masses] * n, **cuts)
    assert np.array_equal(mask.cpu().numpy().astype(bool), (rw != 0).numpy())
    assert torch.allclose(w.cpu(), rw, rtol=1e-8)
    assert torch.allclose(mom.cpu(), rmom, rtol=1e-8, atol=1e-6)


def test_physics_invariants_at_full_size():
    """cfg3 size (2^24 events would need 3 GiB of momenta; 2^22 keeps the check light): 4-momentum
    conservation, on-shell masses, constant massless weight, empty and ragged batches."""
    B = 1 << 22
    r = torch.rand(B, 8, device="cuda", dtype=torch.float64)
    ps = FlatInvertiblePhasespace([100.0] * 2, [100.0] * 4)
    mom, w = ps.generateKinematics_batch(1000.0, r)
    assert float((mom[:, :2].sum(1) - mom[:, 2:].sum(1)).abs().max()) < 5e-9
    mass = torch.sqrt(mom[:, 2:, 0] ** 2 - (mom[:, 2:, 1:] ** 2).sum(-1))
    assert float((mass - 100.0).abs().max()) < 1e-6
    assert bool((w > 0).all())
    ps0 = FlatInvertiblePhasespace([0.0] * 2, [0.0] * 4)
    _, w0 = ps0.generateKinematics_batch(1000.0, r[: 1 << 20])
    assert torch.allclose(w0, torch.full_like(w0, 0.06648282151394422), rtol=1e-12)
    for b in (0, 1, 127, 129):
        m_, w_ = ps.generateKinematics_batch(1000.0, r[:b])
        assert m_.shape == (b, 6, 4) and torch.equal(w_, w[:b])


def test_cut_quirks_of_the_reference():
    """strict '<' comparisons: defaults (-1) and 0 disable the cuts; |max eta| not max |eta|."""
    r = torch.rand(1 << 14, 8, device="cuda", dtype=torch.float64)
    ps = FlatInvertiblePhasespace([0.0] * 2, [0.0] * 4)
    _, w_def = ps.generateKinematics_batch(1000.0, r)
    _, w_zero = ps.generateKinematics_batch(1000.0, r, pT_mincut=0, delR_mincut=0, rap_maxcut=-1)
    assert torch.equal(w_def, w_zero) and bool((w_def > 0).all())
    mom, w_rap = ps.generateKinematics_batch(1000.0, r, rap_maxcut=1.0)
    p = mom[:, 2:]
    eta = torch.asinh(p[..., 3] / torch.sqrt(p[..., 1] ** 2 + p[..., 2] ** 2))
    expect = ~(1.0 < eta.max(1).values.abs())
    assert torch.equal(w_rap != 0, expect)
    assert bool(((eta.abs().max(1).values > 1.0) & (w_rap != 0)).any())     # very negative eta passes, like the reference


@pytest.mark.parametrize("masses,E", [(100.0, 1000.0), (0.0, 1000.0)])
def test_cut_masks_bit_exact_at_a_million_events(masses, E):
    """Cut masks over 2^20 events against the float64 oracle: bit-exact (the kernel decides the cuts on squared
    quantities with series bounds and runs the reference's own formula only inside a thin shell around a threshold)."""
    B = 1 << 20
    r = torch.rand(B, 8, generator=torch.Generator().manual_seed(17), dtype=torch.float64)
    cuts = dict(pT_mincut=20, delR_mincut=0.4, rap_maxcut=2.5)
    ps = FlatInvertiblePhasespace([masses] * 2, [masses] * 4)
    _, w, mask = ps.generateKinematics_batch(E, r.cuda(), return_cutmask=True, **cuts)
    _, rw = orambo.generate_kinematics(E, r, [masses] * 2, [masses] * 4, **cuts)
    assert np.array_equal(mask.cpu().numpy().astype(bool), (rw != 0).numpy())
    assert 0.3 < float(mask.float().mean()) < 0.99


# ---- inverse phase space (the reference has none: README.md:68-69) ---------------------------------------
@pytest.mark.parametrize("case", RAMBO_CASES)
@pytest.mark.parametrize("where", ["cuda", "cpu"])
def test_inverse_recovers_the_reference_uniforms(golden, case, where):
    """nis_rambo_invert on the momenta the REFERENCE produced returns the uniforms the reference was given and, where no
    cut removed the event, the reference's weight; it agrees with the float64 oracle inverse."""
    g = golden("rambo_" + case)
    m = g.meta
    ps = FlatInvertiblePhasespace(m["initial"], m["final"], pdf=None, pdf_active=False)
    r, w = ps.invertKinematics_batch(m["E_cm"], g.t("momenta").to(where))
    assert r.device.type == where and r.dtype == torch.float64 and r.shape == g.t("r").shape
    assert torch.allclose(r.cpu(), g.t("r"), rtol=0, atol=2e-9)
    kept = g.t("weight") != 0
    assert torch.allclose(w.cpu()[kept], g.t("weight")[kept], rtol=1e-7)
    ro, wo = orambo.invert_kinematics(m["E_cm"], g.t("momenta"), m["initial"], m["final"])
    assert torch.allclose(r.cpu(), ro, rtol=0, atol=1e-10) and torch.allclose(w.cpu(), wo, rtol=1e-8)


@pytest.mark.parametrize("masses", [[100.0] * 4, [0.0] * 4, [0.0, 0.0], [10.0, 20.0, 30.0], [0.0, 5.0, 0.0, 80.0, 1.0],
                                    [1.0] * 8], ids=["m4", "m0_4", "n2", "mixed3", "mixed5", "n8"])
def test_inverse_round_trip_at_size(masses):
    """generate -> invert -> generate on 2^18 events for every multiplicity 2..8: the uniforms come back to 1e-9 (the
    azimuth modulo 1), the weight to 1e-9, and regenerating from the recovered uniforms reproduces the momenta."""
    n = len(masses)
    ps = FlatInvertiblePhasespace([0.0, 0.0], masses)
    gen = torch.Generator(device="cuda").manual_seed(77)
    r = torch.rand(1 << 18, 3 * n - 4, device="cuda", dtype=torch.float64, generator=gen)
    mom, w = ps.generateKinematics_batch(1000.0, r)
    r2, w2 = ps.invertKinematics_batch(1000.0, mom)
    d = (r2 - r).abs()
    d = torch.minimum(d, 1.0 - d)                              # phi = 0 and phi = 1 are the same direction
    # a uniform is recovered through K_{j+1} / K_j and a polynomial with slope e (e + 1) u^(e-1) (1 - u): 1e-9 leaves room
    # for the cancellation in M = sqrt(Q^2) when the remaining system is light
    assert float(torch.quantile(d.reshape(-1)[:: 7], 0.9999)) < 1e-10 and float(d.max()) < 1e-6, float(d.max())
    assert torch.allclose(w2, w, rtol=1e-7)
    mom2, _ = ps.generateKinematics_batch(1000.0, r2)
    assert torch.allclose(mom2, mom, rtol=1e-7, atol=1e-7 * 1000.0)


def test_inverse_argument_errors():
    ps = FlatInvertiblePhasespace([0.0, 0.0], [100.0] * 4)
    with pytest.raises(AssertionError):
        ps.invertKinematics_batch(1000.0, torch.zeros(5, 4, 4, dtype=torch.float64, device="cuda"))
    r, w = ps.invertKinematics_batch(1000.0, torch.zeros(0, 6, 4, dtype=torch.float64, device="cuda"))
    assert r.shape == (0, 8) and w.shape == (0,)

