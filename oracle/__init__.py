"""CPU oracle for the NIS hot path of NGoetz/NF (``nisrep``).

TEST INFRASTRUCTURE ONLY.  This package is a closed-form, float64, CPU restatement of the
reference's algorithm (coupling-cell flows, RAMBO-on-diet, the integrate / variance-loss host
formulas).  It exists so that the CUDA path can be checked against something that runs without
``/root/reference``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  Nothing under ``nf_b200/`` imports it, and
the product never falls back to it.

Parity pin: the reference ships no tests, golden vectors or known-answer fixtures, so
the oracle is pinned against outputs of the reference itself, run in the build container by
``tests/golden/make_golden.py`` (committed) which imports ``/root/reference`` and dumps the fixtures
in ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` holds the oracle to those vectors at
<= 1e-12.
"""
