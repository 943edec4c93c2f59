"""Coupling cells with the class names / constructor arguments / state_dict schema of
nisrep/normalizing_flows/layers/coupling_cells.py:73-254.

A cell module only *holds* parameters: ``self.NN`` is the same ``torch.nn.Sequential`` stack as the
reference's conditioner (BatchNorm1d, Linear(no bias), BatchNorm1d, ReLU, ..., Linear(bias), Reshape) so
that ``state_dict()`` keys are ``NN.{idx}.{weight|bias|running_mean|running_var|num_batches_tracked}``
and reference checkpoints load unchanged.  The arithmetic — conditioner MLP, per-bin softmax,
piecewise-linear / piecewise-quadratic CDF, bin search, Jacobian product — runs in the fused CUDA
kernels: ``forward`` of a single cell is a one-cell flow through the C ABI (nis_flow_forward).
"""
import torch

from .layers import Reshape


def conditioner_stack(pass_through_size, sizes, reshape, hidden_bias=False):
    """coupling_cells.py:84-104 / :230-254: BN(P) -> [Linear(no bias) -> BN -> ReLU] * depth ->
    Linear(bias) -> Reshape(T, K).  ``sizes`` = hidden widths + [T*K].  ``hidden_bias``: the affine cell's hidden
    Linear layers carry a bias (coupling_cells.py:27-38 uses the torch default)."""
    mods = [torch.nn.BatchNorm1d(pass_through_size)]
    fan_in = pass_through_size
    for width in sizes[:-1]:
        mods += [torch.nn.Linear(fan_in, width, bias=hidden_bias), torch.nn.BatchNorm1d(width), torch.nn.ReLU()]
        fan_in = width
    mods += [torch.nn.Linear(fan_in, sizes[-1]), Reshape(reshape[0], reshape[1])]
    return torch.nn.Sequential(*mods)


class RectNN(torch.nn.Module):
    """coupling_cells.py:230-254 — the rectangular conditioner network."""

    def __init__(self, pass_through_size, sizes, reshape):
        super().__init__()
        self.NN = conditioner_stack(pass_through_size, sizes, reshape)


class _CouplingCell(torch.nn.Module):
    kind = None            # "lin" | "quad"

    def __init__(self, flow_size, pass_through_size, n_bins, NN_layers):
        super().__init__()
        self.pass_through_size = pass_through_size
        self.flow_size = flow_size
        self.transform_size = flow_size - pass_through_size
        self.n_bins = n_bins
        self.hidden = list(NN_layers)
        K = self.outputs_per_dim()
        self.NN = conditioner_stack(pass_through_size, self.hidden + [self.transform_size * K],
                                    (self.transform_size, K))
        self._solo = None

    def outputs_per_dim(self):
        raise NotImplementedError

    def forward(self, x):
        """x: [B, flow_size+1] -> [B, flow_size+1]: first ``pass_through_size`` columns condition, the
        rest are transformed, last column accumulates the Jacobian.  One-cell fused flow."""
        from ...flowspec import FlowSpec, flow_apply
        if self._solo is None:
            self._solo = FlowSpec([("0", self)], self.flow_size)
        return flow_apply(self._solo, x, self.training)


class PWLin(_CouplingCell):
    """coupling_cells.py:73-142 — piecewise-linear coupling (n_bins logits per transformed dim)."""
    kind = "lin"

    def outputs_per_dim(self):
        return self.n_bins


class PWQuad(_CouplingCell):
    """coupling_cells.py:144-228 — piecewise-quadratic coupling (n_bins+1 vertex heights followed by
    n_bins bin widths per transformed dim)."""
    kind = "quad"

    def outputs_per_dim(self):
        return 2 * self.n_bins + 1


class AffineCoupling(_CouplingCell):
    """coupling_cells.py:6-70 — affine coupling squashed back into the unit interval: the conditioner returns
    (Z0, Z1) per transformed dimension (``Reshape(2, T)``), y = atan(20 e^{Z0} x + relu(Z1)) / (pi/2), and the Jacobian
    takes 20 e^{Z0} / (v^2 + 1) per dimension and 1/(pi/2) once per cell.  Hidden Linear layers have a bias."""
    kind = "affine"

    def __init__(self, flow_size, pass_through_size, NN_layers):
        torch.nn.Module.__init__(self)
        self.pass_through_size = pass_through_size
        self.flow_size = flow_size
        self.transform_size = flow_size - pass_through_size
        self.n_bins = 1
        self.hidden = list(NN_layers)
        self.NN = conditioner_stack(pass_through_size, self.hidden + [2 * self.transform_size],
                                    (2, self.transform_size), hidden_bias=True)
        self._solo = None

    def outputs_per_dim(self):
        return 2

