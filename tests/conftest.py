import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden:
    """One tests/golden/*.npz fixture (written by tests/golden/make_golden.py from the reference)."""

    def __init__(self, fname):
        self.z = np.load(os.path.join(GOLDEN, fname))
        self.meta = json.loads(str(self.z["meta"])) if "meta" in self.z.files else {}
        if self.meta.get("kind") == "affine":
            self.meta.setdefault("n_bins", 1)            # (no bins; the kernels' descriptor carries 1)

    def __getitem__(self, k):
        return self.z[k]

    def keys(self, prefix=""):
        return [k for k in self.z.files if k.startswith(prefix)]

    def t(self, k, dtype=torch.float64):
        return torch.from_numpy(np.asarray(self.z[k])).to(dtype)

    def state_dict(self, dtype=torch.float64):
        sd = {}
        for k in self.keys("sd/"):
            v = torch.from_numpy(np.asarray(self.z[k]))
            sd[k[3:]] = v.to(dtype) if v.dtype.is_floating_point else v
        return sd


FLOW_CASES = ["quad2d", "quad3d", "quad7d", "quad8d", "quad8d_small", "quad9d_extra", "quad16d",
              "lin8d", "lin4d", "lin5d", "affine4d", "affine6d"]
GRAD_CASES = ["quad2d", "quad3d", "quad8d_small", "lin4d", "affine4d"]
RAMBO_CASES = ["m4_cuts", "m4_nocuts", "m0_4", "m2", "mixed3", "m5_cuts", "m0_6", "readme"]
# uniforms on the ends of [0,1] (0, denormal, 2^-24, 1-2^-24, 1) in every column; edge0 = with 0 / denormal rows
# (the reference then stops its lattice bisection after 60 levels), edge1 = without
RAMBO_EDGE_CASES = [t + c for t in ("edge0_", "edge1_") for c in ("m0_4", "m4", "m4_cuts", "m0_4_cuts", "m5", "m0_3")]


# pdf-active phase space (tau / y_cm or x1, x2 sampling, per-event E_cm, lab boost before the cuts), dumped from the
# reference with the analytic stand-in PDF of tests/pdf_stub.py
RAMBO_PDF_CASES = ["pdf_tau_m0_3", "pdf_tau_m4", "pdf_x_m0_4", "pdf_x_m3_nocuts", "pdf_tau_nopdf"]


def rambo_edge_rows(ref_mom, r, n_final):
    """Rows of an edge fixture on which momenta / cut decisions are comparable with the reference's.  The
    reference boosts with gamma = 1/sqrt(1 - beta^2) (utils.py:66-81), which loses eps * gamma^2 and overflows for a
    parent of mass ~2^-30 K: after a mass-dimension uniform of 0 / denormal its momenta are inf / NaN, after 2^-24
    they are off by 1e-8 .. O(1) relative (checked against 50-digit arithmetic: the product's boost by the on-shell
    Q stays at 1e-13), and its cut comparisons on NaNs pass everything.  So momenta and masks are compared where
    every mass-dimension uniform is >= 1e-3 (gamma <~ 10); weights (which do not involve the boost) everywhere."""
    fin = np.isfinite(np.asarray(ref_mom)).all(-1).all(-1)
    return fin & (np.asarray(r)[:, :n_final - 2] >= 1e-3).all(1)


def rambo_edge_weight_rtol(r, meta):
    """Per-row tolerance of the weight on an edge fixture: 1e-9 + 32 eps x the condition number of the subtractions in
    the reference's own formula (flat_phase_space_generator.py:107-113, 394-403).  With K_{j+1} = sqrt(u_j) K_j,
    M_j = K_j + sum_{i>=j} m_i the reweighting evaluates M_j^2 - (M_{j+1} + m_j)^2 = (K_j - K_{j+1})(...) and
    K_j^2 - K_{j+1}^2 by squaring and subtracting: for u_j = 1 - 2^-27/e (r = 1) or K_j ~ 2^-30 (after r = 0) the
    difference is 1e-9 .. 1e-16 of the squares, so any two float64 evaluations (the reference's ATen ops, the
    product's fused multiply-adds) agree only to eps x that ratio.  Rows whose bound exceeds 1e-2 hold rounding
    noise in the reference itself (inf tolerance: only finiteness is asserted there)."""
    from oracle import rambo as orambo
    r = torch.as_tensor(np.asarray(r), dtype=torch.float64)
    m = torch.tensor(meta["final"], dtype=torch.float64)
    n = len(meta["final"])
    K = torch.zeros(r.shape[0], n - 1, dtype=torch.float64)
    K[:, 0] = meta["E_cm"] - m.sum()
    if n > 2:
        u = orambo.bisect(r[:, :n - 2].clone(), n)
        for i in range(2, n):
            K[:, i - 1] = torch.sqrt(u[:, i - 2] * K[:, i - 2] ** 2)
    msum = torch.flip(torch.cumsum(torch.flip(m, (-1,)), -1), (-1,))
    M = torch.cat((K + msum[:-1], m[-1:].unsqueeze(0).repeat(r.shape[0], 1)), -1)
    amp = torch.ones(r.shape[0], dtype=torch.float64)
    for j in range(n - 1):
        amp = torch.maximum(amp, M[:, j] ** 2 / (M[:, j] ** 2 - (M[:, j + 1] + m[j]) ** 2).abs().clamp_min(1e-300))
        if j < n - 2:
            amp = torch.maximum(amp, K[:, j] ** 2 / (K[:, j] ** 2 - K[:, j + 1] ** 2).abs().clamp_min(1e-300))
    tol = 1e-9 + 32 * 2.220446049250313e-16 * amp
    return torch.where(tol > 1e-2, torch.full_like(tol, float("inf")), tol).numpy()


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name + ".npz")
        return cache[name]
    return get
