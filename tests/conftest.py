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
              "lin8d", "lin4d", "lin5d"]
GRAD_CASES = ["quad2d", "quad3d", "quad8d_small", "lin4d"]
RAMBO_CASES = ["m4_cuts", "m4_nocuts", "m0_4", "m2", "mixed3", "m5_cuts", "m0_6", "readme"]
# uniforms on the ends of [0,1] (0, denormal, 2^-24, 1-2^-24, 1) in every column; edge0 = with 0 / denormal rows
# (the reference then stops its lattice bisection after 60 levels), edge1 = without
RAMBO_EDGE_CASES = [t + c for t in ("edge0_", "edge1_") for c in ("m0_4", "m4", "m4_cuts", "m0_4_cuts", "m5", "m0_3")]


def rambo_edge_rows(ref_mom, r, n_final):
    """Rows of an edge fixture on which momenta / cut decisions are comparable with the reference's.  The
    reference boosts with gamma = 1/sqrt(1 - beta^2) (utils.py:66-81), which loses eps * gamma^2 and overflows for a
    parent of mass ~2^-30 K: after a mass-dimension uniform of 0 / denormal its momenta are inf / NaN, after 2^-24
    they are off by 1e-8 .. O(1) relative (checked against 50-digit arithmetic: the product's boost by the on-shell
    Q stays at 1e-13), and its cut comparisons on NaNs pass everything.  So momenta and masks are compared where
    every mass-dimension uniform is >= 1e-3 (gamma <~ 10); weights (which do not involve the boost) everywhere."""
    fin = np.isfinite(np.asarray(ref_mom)).all(-1).all(-1)
    return fin & (np.asarray(r)[:, :n_final - 2] >= 1e-3).all(1)


def rambo_edge_weight_rtol(r, n_final):
    """Per-row tolerance of the weight on an edge fixture: 1e-9, except 1e-6 where a mass-dimension uniform is 1:
    there u = 1 - 2^-27/e and the reweighting factor 1/(K_j^2 - K_{j+1}^2) = 1/(K_j^2 (1-u)) turns the one rounding of
    K_{j+1} = sqrt(u) K_j into 1e-16 / 4e-9 ~ 3e-8 (in the reference as much as in the product)."""
    return np.where((np.asarray(r)[:, :n_final - 2] >= 1.0).any(1), 1e-6, 1e-9)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name + ".npz")
        return cache[name]
    return get
