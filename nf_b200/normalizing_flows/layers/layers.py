"""Permutation / bookkeeping layers with the class names, constructor arguments and public attributes
of nisrep/normalizing_flows/layers/layers.py:6-91.

Design: every coordinate-shuffling layer is a ``ColumnGather`` — it only knows how to produce a gather
index over the ``n_flow`` coordinate columns (``gather_index(d)``).  ``nf_b200.flowspec.FlowSpec``
composes those indices on the host into per-cell column tables, so inside a ``FlowSequential`` no data
ever moves.  Called on its own a layer applies its index once (``tensor[:, idx]``).
"""
import torch


class ColumnGather(torch.nn.Module):
    """out[:, i] = in[:, gather_index(d)[i]] for the d coordinate columns; the Jacobian column stays."""

    def gather_index(self, d):
        raise NotImplementedError

    def forward(self, tensor):
        d = tensor.shape[-1] - 1
        idx = torch.as_tensor(self.gather_index(d) + [d], dtype=torch.long, device=tensor.device)
        return tensor.index_select(-1, idx)


class RollLayer(ColumnGather):
    """layers.py:80-91 — cyclic shift: out[:, (i + shift) % d] = in[:, i]."""

    def __init__(self, shift):
        super().__init__()
        self.shift = shift

    def gather_index(self, d):
        return [(i - self.shift) % d for i in range(d)]


class MaskLayer(ColumnGather):
    """layers.py:6-32 — cell number ``pos`` looks at binary digit ``pos // 2`` (MSB first) of every
    dimension index: dims whose digit equals ``pos % 2`` feed the conditioner (``feeder``), the others
    are transformed (``trafoer``); output order is feeder, trafoer."""

    def __init__(self, dims_bin, pos, dev):
        super().__init__()
        digit = torch.as_tensor(dims_bin)[:, pos // 2]
        want = pos % 2
        self.feeder = (digit == want).nonzero().to(dev)        # [P, 1] like the reference
        self.trafoer = (digit != want).nonzero().to(dev)       # [T, 1]
        self.pass_through = self.feeder.shape[0]

    def gather_index(self, d):
        return self.feeder.view(-1).tolist() + self.trafoer.view(-1).tolist()


class DeMaskLayer(ColumnGather):
    """layers.py:34-51 — inverse of the MaskLayer built from the same (feeder, trafoer)."""

    def __init__(self, first, second):
        super().__init__()
        self.list_ind = torch.cat((first, second), 0).view(1, -1)

    def gather_index(self, d):
        fwd = self.list_ind.view(-1).tolist()
        inv = [0] * len(fwd)
        for pos, col in enumerate(fwd):
            inv[col] = pos
        return inv


class Reshape(torch.nn.Module):
    """layers.py:55-64 — [B, T*K] -> [B, T, K] (K fastest), a copy like the reference's clone()."""

    def __init__(self, shapes1, shapes2):
        super().__init__()
        self.shapes = (shapes1, shapes2)

    def forward(self, tensor):
        return tensor.reshape(tensor.shape[0], *self.shapes).clone()


class AddJacobian(torch.nn.Module):
    """layers.py:66-77 — append the Jacobian column (float64 like the reference, which makes the whole
    [B, n_flow+1] tensor float64)."""

    def __init__(self, jacobian_value=torch.ones(1)):
        super().__init__()
        self.jacobian_value = jacobian_value

    def forward(self, input, dev=torch.device("cpu")):
        x = input.to(dev)
        # filled on the device (no host-to-device copy per call: that would also break CUDA-graph capture of the epoch)
        j = torch.full((x.shape[0], 1), float(self.jacobian_value), dtype=torch.double, device=x.device)
        return torch.cat((x, j), dim=1)
