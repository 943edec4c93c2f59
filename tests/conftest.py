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


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = Golden(name + ".npz")
        return cache[name]
    return get
