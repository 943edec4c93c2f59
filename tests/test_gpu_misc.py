"""GPU checks of the reduction / RNG kernels and of integrate() against analytic answers."""
import ctypes
import math

import pytest
import torch

from nf_b200 import _cabi
from nf_b200.normalizing_flows.manager import PWQuadManager

pytestmark = pytest.mark.gpu


def camel(x):
    return torch.exp(-((x[:, 0] - 0.75) ** 2 + (x[:, 1] - 0.75) ** 2) / (0.2 ** 2)) + \
        torch.exp(-((x[:, 0] - 0.25) ** 2 + (x[:, 1] - 0.25) ** 2) / (0.2 ** 2))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n", [0, 1, 255, 1 << 20, (1 << 22) + 3])
def test_reduce_moments(dtype, n):
    lib = _cabi.lib()
    v = torch.randn(n, device="cuda", dtype=dtype) + 3
    out = torch.full((3,), 7.0, device="cuda", dtype=torch.double)
    ws = torch.empty(lib.nis_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    for acc in (0, 1):
        _cabi.check(lib.nis_reduce_moments(_cabi.ptr(v), _cabi.dtype_code(v), n, _cabi.ptr(out), acc, _cabi.ptr(ws),
                                           ws.numel(), _cabi.stream_ptr()), "reduce")
    ref = torch.stack((v.double().sum(), (v.double() ** 2).sum(), torch.tensor(float(n), device="cuda", dtype=torch.double)))
    assert torch.allclose(out, 2 * ref, rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("n", [1, 255, (1 << 21) + 3])
def test_reduce_stats_matches_torch(dtype, n):
    """nis_reduce_stats: (sum, sum of squares, n, max, min, #non-finite) - the fused reduction behind integrate()
    and the unweighting statistics (experiment_mg.py:73-76) against torch on the same vector."""
    lib = _cabi.lib()
    ws = torch.empty(lib.nis_reduce_workspace_bytes(), dtype=torch.uint8, device="cuda")
    v = torch.randn(n, device="cuda", dtype=dtype) * 3 + 1
    out = torch.zeros(6, device="cuda", dtype=torch.double)

    def run(acc):
        _cabi.check(lib.nis_reduce_stats(_cabi.ptr(v), _cabi.dtype_code(v), n, _cabi.ptr(out), acc, _cabi.ptr(ws),
                                         ws.numel(), _cabi.stream_ptr()), "stats")
    run(0)
    vd = v.double()
    ref = torch.stack((vd.sum(), (vd ** 2).sum(), torch.tensor(float(n), device="cuda", dtype=torch.double),
                       vd.max(), vd.min(), torch.zeros((), device="cuda", dtype=torch.double)))
    assert torch.allclose(out, ref, rtol=1e-12, atol=1e-9) and float(out[3]) == float(vd.max())
    run(1)                                                  # accumulate: sums double, extremes stay
    assert torch.allclose(out[:3], 2 * ref[:3], rtol=1e-12, atol=1e-9) and torch.equal(out[3:5], ref[3:5])
    if n > 10:                                              # non-finite entries are counted and propagate like torch's
        v[3] = float("inf")
        v[7] = float("nan")
        v[n - 1] = -float("inf")
        run(0)
        assert float(out[5]) == 3.0 and torch.isnan(out[0]) and torch.isnan(out[3]) and torch.isnan(out[4])
        v[7] = 0.0
        run(0)
        assert float(out[5]) == 2.0 and float(out[3]) == float("inf") and float(out[4]) == -float("inf")


def test_uniform_fill_is_counter_based():
    lib = _cabi.lib()
    n = 1 << 20
    for dt, code in ((torch.float32, _cabi.F32), (torch.float64, _cabi.F64)):
        a = torch.empty(n, device="cuda", dtype=dt)
        b = torch.empty(n // 2, device="cuda", dtype=dt)
        _cabi.check(lib.nis_uniform_fill(_cabi.ptr(a), code, n, 42, 0, _cabi.stream_ptr()), "fill")
        _cabi.check(lib.nis_uniform_fill(_cabi.ptr(b), code, n // 2, 42, n // 2, _cabi.stream_ptr()), "fill")
        assert torch.equal(a[n // 2:], b)                       # element i depends only on (seed, offset+i)
        assert float(a.min()) >= 0.0 and float(a.max()) < 1.0
        assert abs(float(a.double().mean()) - 0.5) < 5 / math.sqrt(12 * n)
        assert abs(float(a.double().var()) - 1 / 12) < 1e-3
        c = torch.empty(n, device="cuda", dtype=dt)
        _cabi.check(lib.nis_uniform_fill(_cabi.ptr(c), code, n, 43, 0, _cabi.stream_ptr()), "fill")
        assert not torch.equal(a, c)


def test_integrate_untrained_flow_reproduces_the_analytic_camel_integral(golden):
    """Integral estimates must agree with the reference within one combined standard error; the
    reference's own numbers for this very model (same state_dict) are in the golden file."""
    g = golden("integrate_camel")
    analytic = 2 * (0.5 * math.sqrt(0.04 * math.pi) * (math.erf(3.75) + math.erf(1.25))) ** 2
    torch.manual_seed(123)
    NF = PWQuadManager(n_flow=2)
    NF.create_model(2, 4, [3] * 3)
    NF._model.load_state_dict(g.state_dict())
    for train, key in ((True, ""), (False, "_eval")):
        NF.best_model.train(train)
        sig, err = NF.integrate(camel, 10, 10000, 0)
        honest = float(err) * math.sqrt(10)                     # manager.py:403 under-reports by sqrt(nitn)
        ref_sig, ref_err = float(g["sig" + key]), float(g["err" + key]) * math.sqrt(10)
        assert abs(float(sig) - ref_sig) < 3 * math.hypot(honest, ref_err)
        assert abs(float(sig) - analytic) < 4 * honest
        assert 0.3 < float(err) / float(g["err" + key]) < 3.0


def test_end_to_end_integrate_flow_rambo_known_answer():
    """BASELINE configs[3] through the public API at 2^24 points: 8-D PWQuad flow (6 mask cells, 32 bins, [64]*3)
    -> RAMBO 2->4 massless at E_cm = 1000 -> |M|^2 = 1.  The flow is a bijection of the unit cube, so the estimate
    must be the flat-weight constant 0.06648282151394422 within ONE combined honest standard error (the reference's
    value is that constant exactly: its massless weight does not depend on the point).  float32 uniforms hit exactly
    0 with p = 2^-24 per coordinate - 2^24 x 8 draws hold ~8 of them - which used to turn the estimate into NaN."""
    import warnings
    from nf_b200.PhaseSpace.flat_phase_space_generator import FlatInvertiblePhasespace
    torch.manual_seed(1234)
    NF = PWQuadManager(n_flow=8)
    NF.create_model(6, 32, [64] * 3)
    ps = FlatInvertiblePhasespace([0.0] * 2, [0.0] * 4)
    with warnings.catch_warnings():
        warnings.simplefilter("error")                         # a NonFiniteWeightWarning fails the test
        sig, err = NF.integrate(lambda X: ps.generateKinematics_batch(1000.0, X, momenta=False)[1], 4, 1 << 22, 0)
    assert NF.n_nonfinite == 0
    honest = float(err) * math.sqrt(4)
    assert math.isfinite(float(sig)) and abs(float(sig) - 0.06648282151394422) <= honest, (float(sig), honest)
    st = NF.weight_statistics(lambda X: ps.generateKinematics_batch(1000.0, X, momenta=False)[1], 1 << 22, 0)
    assert st["n_nonfinite"] == 0 and st["n"] == 1 << 22 and 0 < st["unweighting_efficiency"] <= 1
    assert abs(st["w_mean"] - 0.06648282151394422) <= 3 * math.sqrt(st["v_var"] / st["n"])


def test_integrate_reports_non_finite_weights():
    from nf_b200.normalizing_flows.manager import NonFiniteWeightWarning
    torch.manual_seed(3)
    NF = PWQuadManager(n_flow=2)
    NF.create_model(2, 4, [3] * 3)

    def bad(x):
        y = camel(x)
        y[5] = float("nan")
        return y
    with pytest.warns(NonFiniteWeightWarning):
        sig, _ = NF.integrate(bad, 2, 1000, 0)
    assert NF.n_nonfinite == 2 and not math.isfinite(float(sig))


def test_weight_statistics_match_torch():
    """experiment_mg.py:66-76,101: v_var / w_max / w_mean of f(x) J(x) over one batch through best_model."""
    torch.manual_seed(4)
    NF = PWQuadManager(n_flow=2)
    NF.create_model(2, 4, [3] * 3)
    seen = {}

    def f(x):
        seen["x"] = x
        return camel(x)
    st = NF.weight_statistics(f, 5000, 0)
    with torch.no_grad():
        # the same points again: x -> weights by hand
        X = seen["x"]
    assert X.shape == (5000, 2)
    # recompute from the latent points is not possible (the flow consumed them); check internal consistency instead
    assert st["w_min"] <= st["w_mean"] <= st["w_max"] and st["v_var"] > 0
    assert abs(st["unweighting_efficiency"] - st["w_mean"] / st["w_max"]) < 1e-15


def test_train_mode_batch_of_one_is_refused_like_torch():
    torch.manual_seed(5)
    NF = PWQuadManager(n_flow=2)
    NF.create_model(2, 4, [3] * 3)
    NF._model.train()
    with pytest.raises(ValueError, match="more than 1 value per channel"):
        NF._model(torch.rand(1, 2, device="cuda", dtype=torch.float64))
    NF._model.eval()
    assert NF._model(torch.rand(1, 2, device="cuda", dtype=torch.float64)).shape == (1, 3)
