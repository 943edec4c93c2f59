"""nf_b200 — B200-native (sm_100a) implementation of the neural-importance-sampling hot path of NGoetz/NF.

Same Python surface as the reference package ``nisrep``:

    from nf_b200.normalizing_flows.manager import PWQuadManager, PWLinManager
    from nf_b200.PhaseSpace.flat_phase_space_generator import FlatInvertiblePhasespace

PyTorch is the host layer (device memory, streams, autograd plumbing, torch.distributed); all arithmetic
on the hot path runs in hand-written CUDA kernels reached through the C ABI of ``libnisb200.so``
(include/nis_b200.h).  There is no CPU fallback.
"""
__version__ = "0.1"
