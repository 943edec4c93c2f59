"""A stand-in for an LHAPDF set, used to pin the pdf-active phase-space path (LHAPDF is not in this image and not in
the reference tree).  Same call the reference makes — ``pdf.xfxQ2(pdg, x, Q2)`` with tensors
(flat_phase_space_generator.py:133) — returning x f(x) = A x^a (1-x)^b per parton, Q^2-independent."""
import numpy as np
import torch

SHAPES = {21: (3.0, -0.2, 5.0), 1: (1.1, 0.6, 4.0), 2: (2.0, 0.5, 3.0), -1: (0.3, -0.1, 7.0), -2: (0.25, -0.1, 7.0)}


class StubPdf:
    def xfxQ2(self, pdg, x, q2):
        x = x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x, dtype=np.float64)
        A, a, b = SHAPES.get(pdg, (1.0, 0.0, 3.0))
        return (A * x ** a * (1.0 - x) ** b).tolist()
