"""RAMBO-on-diet flat phase-space generator — the public surface of
nisrep/PhaseSpace/flat_phase_space_generator.py (VirtualPhaseSpaceGenerator, FlatInvertiblePhasespace,
PhaseSpaceGeneratorError) on top of the fused sm_100a kernel ``nis_rambo_generate``: intermediate
masses, massive reweighting, sequential two-body decays with boosts, the pT / deltaR / rapidity cuts and
the 1/(2 s) flux factor in ONE pass per event (the reference: a 120-180 level batched bisection plus
~600 ATen launches).

Both modes of the reference: ``pdf_active=False`` (what BASELINE.json names) and ``pdf_active=True`` (two more
uniforms sample the Bjorken x of the beams, per-event partonic energy, parton densities, lab-frame cuts;
flat_phase_space_generator.py:157-187).  The densities come from any object with the reference's
``pdf.xfxQ2(pdg, x, Q2)`` call (an ``lhapdf.PDF``; LHAPDF itself is not needed to construct the generator), sampled
once per parton into a device grid the kernel interpolates (pdf_grid.py).
"""
import ctypes
import math

import torch

from .. import _cabi
from .pdf_grid import PdfGrid, X_CUT, is_parton


class PhaseSpaceGeneratorError(Exception):
    pass


class VirtualPhaseSpaceGenerator(object):
    """flat_phase_space_generator.py:23-54."""

    def __init__(self, initial_masses, final_masses, pdf=None, pdf_active=False, tau=True):
        dev = torch.device("cuda:0") if torch.cuda.is_available() else torch.device("cpu")
        self.initial_masses = initial_masses
        self.masses_t = torch.tensor(final_masses, requires_grad=False, dtype=torch.double, device=dev)
        self.n_initial = len(initial_masses)
        self.n_final = len(final_masses)
        self.pdf = pdf
        self.pdf_active = pdf_active
        self.tau = tau
        self._pdf_grids = {}

    def generateKinematics(self, E_cm, random_variables):
        raise NotImplementedError

    def nDimPhaseSpace(self):
        """Number of uniforms per event: 3 n_final - 4 (0 for a single final particle)."""
        return 0 if self.n_final == 1 else 3 * self.n_final - 4


class FlatInvertiblePhasespace(VirtualPhaseSpaceGenerator):
    """Implementation following S. Platzer, arXiv:1308.2922 (flat_phase_space_generator.py:57-441)."""

    epsilon_border = 1e-10
    absolute_Ecm_min = 1.
    check_nan = True      # the reference scans the uniforms for NaN on every call (:147-149)

    def __init__(self, *args, **opts):
        super(FlatInvertiblePhasespace, self).__init__(*args, **opts)
        if self.n_initial == 1:
            raise PhaseSpaceGeneratorError("This basic generator does not support decay topologies.")
        if self.n_initial > 2:
            raise PhaseSpaceGeneratorError("This basic generator does not support more than 2 incoming particles.")
        if not 2 <= self.n_final <= _cabi.NIS_MAX_FINAL:
            raise PhaseSpaceGeneratorError("This build supports 2 to %d final-state particles." % _cabi.NIS_MAX_FINAL)

    @staticmethod
    def get_flatWeights(E_cm, n):
        """Massless n-body phase-space volume (2 pi)^(4-3n) (pi/2)^(n-1) s^(n-2) / ((n-1)! (n-2)!)
        (flat_phase_space_generator.py:81-97)."""
        if n == 1:
            return 1.
        norm = math.pow(2 * math.pi, 4 - 3 * n) * math.pow(math.pi / 2.0, n - 1) / \
            (math.factorial(n - 1) * math.factorial(n - 2))
        if torch.is_tensor(E_cm):
            return norm * torch.pow(E_cm ** 2, n - 2)
        return norm * math.pow(E_cm ** 2, n - 2)

    def nDimInput(self):
        """Columns of the uniforms handed to generateKinematics_batch: nDimPhaseSpace() + 2 in pdf-active mode (:159)."""
        return self.nDimPhaseSpace() + (2 if self.pdf_active else 0)

    def _pdf_desc(self, d, E_cm, pdgs, dev):
        """Parton-density part of the descriptor (:157-187).  Returns the device grids (kept alive by the caller)."""
        d.pdf_active, d.tau_mode = 1, 1 if self.tau else 0
        tot = float(torch.sum(self.masses_t))
        d.tau_min = (max(tot, self.absolute_Ecm_min) / float(E_cm)) ** 2          # :163-164
        d.x_cut = X_CUT
        keep = []
        for i, pdg in enumerate(pdgs[:2]):
            if self.pdf is None or not is_parton(pdg):                           # get_pdfQ2 :124-128 -> density 1
                d.pdf_grid[i] = None
                continue
            if pdg not in self._pdf_grids:
                self._pdf_grids[pdg] = PdfGrid(self.pdf, pdg)
            g = self._pdf_grids[pdg]
            t = g.on(dev)
            keep.append(t)
            d.pdf_grid[i] = t.data_ptr()
            d.pdf_nodes, d.pdf_lnx_lo = g.n_nodes, g.lnx_lo
        return keep

    def _desc(self, E_cm, pT_mincut, delR_mincut, rap_maxcut):
        d = _cabi.NisRamboDesc()
        d.n_final = self.n_final
        d.initial_masses[0], d.initial_masses[1] = float(self.initial_masses[0]), float(self.initial_masses[1])
        for i, m in enumerate(self.masses_t.tolist()):
            d.final_masses[i] = m
        d.E_cm = float(E_cm)
        d.pT_mincut, d.delR_mincut, d.rap_maxcut = float(pT_mincut), float(delR_mincut), float(rap_maxcut)
        return d

    def generateKinematics_batch(self, E_cm, random_variables_full, pT_mincut=-1, delR_mincut=-1, rap_maxcut=-1,
                                 pdgs=[0, 0], return_cutmask=False, momenta=True):
        """r[B, 3 n_final - 4] uniforms -> (momenta[B, 2+n_final, 4] float64 in the CM frame, weight[B]
        float64 = flat weight x massive Jacobian x cuts / (2 s))  (flat_phase_space_generator.py:139-308).
        In pdf-active mode r has two more columns (tau, y_cm; or x2, x1 with ``tau=False``), ``pdgs`` name the two
        partons, and the weight also carries the sampling Jacobian, both parton densities and the x cut (:157-187).

        Extras over the reference signature: ``return_cutmask`` appends the uint8 pass mask,
        ``momenta=False`` skips writing the momenta (weight-only mode) and returns None for them.
        Results come back on the device of ``random_variables_full`` (the kernel always runs on CUDA).
        """
        r = random_variables_full
        if not torch.is_tensor(r):
            r = torch.as_tensor(r, dtype=torch.double)
        self.collider_energy = E_cm
        if self.check_nan and torch.isnan(r).any():
            raise PhaseSpaceGeneratorError("Some of the random variables passed to the phase-space generator are NaN")
        assert r.dim() == 2 and r.shape[1] == self.nDimInput()
        if torch.is_tensor(E_cm):
            raise TypeError("E_cm is the (scalar) collider energy; the per-event partonic energy sqrt(x1 x2) E_cm "
                            "is formed inside the kernel in pdf-active mode")
        lib = _cabi.lib()
        home = r.device
        dev = home if home.type == "cuda" else self.masses_t.device
        rd = r.detach().to(dev).contiguous()
        if rd.dtype not in (torch.float32, torch.float64):
            rd = rd.double()
        if rd.data_ptr() % 16:                     # the kernel moves rows 16 bytes at a time (a slice can start anywhere)
            rd = rd.clone()
        B = rd.shape[0]
        desc = self._desc(E_cm, pT_mincut, delR_mincut, rap_maxcut)
        grids = self._pdf_desc(desc, E_cm, pdgs, dev) if self.pdf_active else None      # noqa: F841 (keeps the grids alive)
        with torch.cuda.device(dev):
            mom = torch.empty(B, 2 + self.n_final, 4, dtype=torch.double, device=dev) if momenta else None
            weight = torch.empty(B, dtype=torch.double, device=dev)
            mask = torch.empty(B, dtype=torch.uint8, device=dev) if return_cutmask else None
            rc = lib.nis_rambo_generate(ctypes.byref(desc), _cabi.ptr(rd), _cabi.dtype_code(rd), _cabi.ptr(mom),
                                        _cabi.ptr(weight), _cabi.ptr(mask), B, _cabi.stream_ptr(dev))
            _cabi.check(rc, "nis_rambo_generate")
        if home != dev:
            mom = mom.to(home) if mom is not None else None
            weight = weight.to(home)
            mask = mask.to(home) if mask is not None else None
        if return_cutmask:
            return mom, weight, mask
        return mom, weight

    def invertKinematics_batch(self, E_cm, momenta):
        """momenta[B, 2+n_final, 4] (as ``generateKinematics_batch`` returns them: CM frame, two beam rows first) ->
        (random_variables[B, 3 n_final - 4] float64, weight[B] float64): the uniforms the generator maps to these momenta and
        the flat weight x massive Jacobian / (2 s) of that point, WITHOUT cuts.  The reference lists the inverse as to do
        (README.md:68-69): this is the algebraic inverse of :139-308 (masses of the remaining system ->
        K ratios -> the mass polynomial evaluated forward; decay angles in the parent rest frame).  pdf-inactive only.
        """
        if self.pdf_active:
            raise NotImplementedError("invertKinematics_batch: the pdf-active map (tau / y_cm) is not inverted")
        mom = momenta if torch.is_tensor(momenta) else torch.as_tensor(momenta, dtype=torch.double)
        assert mom.dim() == 3 and mom.shape[1] == 2 + self.n_final and mom.shape[2] == 4
        if torch.is_tensor(E_cm):
            raise TypeError("E_cm is the (scalar) centre-of-mass energy")
        lib = _cabi.lib()
        home = mom.device
        dev = home if home.type == "cuda" else self.masses_t.device
        md = mom.detach().to(dev, torch.double).contiguous()
        if md.data_ptr() % 32:                     # events are read 32 bytes at a time
            md = md.clone()
        B = md.shape[0]
        desc = self._desc(E_cm, -1, -1, -1)
        with torch.cuda.device(dev):
            r = torch.empty(B, self.nDimPhaseSpace(), dtype=torch.double, device=dev)
            weight = torch.empty(B, dtype=torch.double, device=dev)
            rc = lib.nis_rambo_invert(ctypes.byref(desc), _cabi.ptr(md), _cabi.ptr(r), _cabi.ptr(weight), B,
                                      _cabi.stream_ptr(dev))
            _cabi.check(rc, "nis_rambo_invert")
        if home != dev:
            r, weight = r.to(home), weight.to(home)
        return r, weight

