"""Device-side parton densities for the pdf-active phase space.

The reference calls ``pdf.xfxQ2(pdg, x, Q2)`` (LHAPDF) per batch at the fixed scale Q^2 = 91.188^2
(flat_phase_space_generator.py:120-137, 184).  LHAPDF is a host library; here the same object is sampled ONCE per
parton on a grid uniform in ln x and the fused RAMBO kernel interpolates it (4-point Lagrange in ln x), so a batch
never leaves the device.  Any object with the reference's ``xfxQ2(pdg, x, Q2)`` call works (an ``lhapdf.PDF`` or the
analytic stand-in used by the tests)."""
import math

import numpy as np
import torch

Q2_REF = 91.188 ** 2          # flat_phase_space_generator.py:184
X_CUT = 1e-4                  # :185-186


def is_parton(pdg):
    """get_pdfQ2 returns 1 unless the code is a gluon or a quark (:127-128)."""
    return pdg in [21] or abs(pdg) in range(1, 7)


class PdfGrid:
    """x f(x, Q2_REF) of one parton on ``n_nodes`` points uniform in ln x over [ln x_lo, 0]."""

    def __init__(self, pdf, pdg, n_nodes=16384, x_lo=0.5 * X_CUT):
        self.pdg, self.n_nodes, self.lnx_lo = int(pdg), int(n_nodes), math.log(x_lo)
        x = torch.exp(torch.linspace(self.lnx_lo, 0.0, self.n_nodes, dtype=torch.float64))
        x[-1] = 1.0
        f = pdf.xfxQ2(self.pdg, x, torch.full_like(x, Q2_REF))
        self.host = torch.as_tensor(np.asarray(f, dtype=np.float64)).reshape(-1).contiguous()
        if self.host.numel() != self.n_nodes:
            raise ValueError("pdf.xfxQ2 must return one value per x")
        self._dev = {}

    def on(self, device):
        key = (device.type, device.index)
        if key not in self._dev:
            self._dev[key] = self.host.to(device)
        return self._dev[key]
