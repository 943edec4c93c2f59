"""Host side of the fused flow: topology folding, parameter arenas, autograd glue, C-ABI calls.

``FlowSpec`` turns an ordered list of reference-style modules (PWLin / PWQuad cells, RollLayer,
MaskLayer, DeMaskLayer — what ``create_model`` registers, manager.py:484-492,538-585 of the reference)
into the ``NisFlowDesc`` of include/nis_b200.h and keeps the cells' parameters in ONE float32 arena (the
``nn.Parameter``s are views into it, so an optimizer step updates the arena in place and no gather is
needed per call) plus one arena for the BatchNorm running statistics.
"""
import ctypes
import os

import torch

from . import _cabi
from .normalizing_flows.layers.layers import ColumnGather


def _cell_tensors(cell):
    """Parameters / buffers of one cell in the order of the C parameter block (include/nis_b200.h)."""
    nn_ = cell.NN
    depth = len(cell.hidden)
    params = [nn_[0].weight, nn_[0].bias]
    bns = [nn_[0]]
    for l in range(depth):
        lin, bn = nn_[1 + 3 * l], nn_[2 + 3 * l]
        params += [lin.weight, bn.weight, bn.bias]
        bns.append(bn)
    out = nn_[1 + 3 * depth]
    params += [out.weight, out.bias]
    return params, bns


def _hidden_biases(cell):
    """Biases of the hidden Linear layers (AffineCoupling only), layer by layer.  They stay outside the kernels' parameter
    block: in front of a BatchNorm a bias only moves the mean, so it is folded into the running mean the kernels see
    (FlowSpec._bn_for_kernel) and gets its gradient from the BatchNorm shift's (FlowSpec._hidden_bias_grads)."""
    nn_ = cell.NN
    return [nn_[1 + 3 * l].bias for l in range(len(cell.hidden)) if nn_[1 + 3 * l].bias is not None]


class _Arena:
    """A flat tensor whose slices back a list of tensors (``t.data`` is re-pointed at its slice)."""

    def __init__(self, tensors, dtype):
        self.tensors = tensors
        self.dtype = dtype
        self.offsets = []
        off = 0
        for t in tensors:
            self.offsets.append(off)
            off += t.numel()
        self.total = off
        self.flat = None

    def _intact(self, device):
        f = self.flat
        if f is None or f.device != device:
            return False
        base, isz = f.data_ptr(), f.element_size()
        for t, off in zip(self.tensors, self.offsets):
            if t.dtype != self.dtype or t.device != device or t.data_ptr() != base + off * isz or \
                    not t.is_contiguous():
                return False
        return True

    def get(self, device):
        """The flat tensor on ``device`` with every member a live view of it.  Members of another
        dtype (e.g. after ``model.double()``) are left alone and copied in on every call."""
        if self._intact(device):
            return self.flat
        flat = torch.empty(self.total, dtype=self.dtype, device=device)
        relink = all(t.dtype == self.dtype for t in self.tensors)
        with torch.no_grad():
            for t, off in zip(self.tensors, self.offsets):
                sl = flat[off:off + t.numel()].view(t.shape)
                sl.copy_(t.detach().to(device))
                if relink:
                    t.data = sl
        self.flat = flat if relink else None
        return flat

    def write_back(self, flat):
        """Slow path (members are not views): copy arena contents back into the members."""
        if self.flat is flat:
            return
        with torch.no_grad():
            for t, off in zip(self.tensors, self.offsets):
                t.copy_(flat[off:off + t.numel()].view(t.shape))


class FlowSpec:
    """Compiled flow: descriptor + arenas.  ``named_layers`` is [(name, module), ...] in call order."""

    def __init__(self, named_layers, n_flow):
        d = int(n_flow)
        if not 2 <= d <= _cabi.NIS_MAX_DIM:
            raise ValueError("n_flow must be in [2, %d]" % _cabi.NIS_MAX_DIM)
        from .normalizing_flows.layers.coupling_cells import _CouplingCell
        cur = list(range(d))                      # cur[logical position] = physical column
        cells = []
        for name, mod in named_layers:
            if isinstance(mod, _CouplingCell):
                P = mod.pass_through_size
                cells.append((name, mod, cur[:P], cur[P:]))
            elif isinstance(mod, ColumnGather):
                g = mod.gather_index(d)
                cur = [cur[i] for i in g]
            else:
                raise TypeError("FlowSpec cannot fold a %s" % type(mod).__name__)
        if not cells:
            raise ValueError("a flow needs at least one coupling cell")
        if len(cells) > _cabi.NIS_MAX_CELLS:
            raise ValueError("at most %d coupling cells" % _cabi.NIS_MAX_CELLS)
        first = cells[0][1]
        for _, c, _, _ in cells:
            if (c.kind, c.n_bins, c.hidden, c.flow_size) != (first.kind, first.n_bins, first.hidden, d):
                raise ValueError("all cells of a flow must share kind, n_bins, hidden widths and flow size")
        if len(first.hidden) > _cabi.NIS_MAX_HIDDEN or any(h > _cabi.NIS_MAX_WIDTH for h in first.hidden):
            raise ValueError("conditioner too large for this build")
        self.n_flow = d
        self.kind = first.kind
        self.n_bins = first.n_bins
        self.hidden = list(first.hidden)
        self.cells = cells
        self.out_perm = cur
        self.n_cells = len(cells)

        params, bns = [], []
        self.param_names = []
        for name, c, _, _ in cells:
            p, b = _cell_tensors(c)
            params += p
            bns += b
        self.params = params
        self.bn_modules = bns
        self.param_arena = _Arena(params, torch.float32)
        rstats = []
        for bn in bns:
            rstats += [bn.running_mean, bn.running_var]
        self.bn_arena = _Arena(rstats, torch.float32)
        # hidden-layer biases (affine cells): where the running mean / variance of the BatchNorm behind each of them sits in
        # the bn arena, and where that BatchNorm's weight / bias sit in the parameter arena
        self.hidden_biases = []
        hb_mean, hb_var, hb_gamma, hb_beta = [], [], [], []
        bn_i = 0
        for _, c, _, _ in cells:
            hbs = _hidden_biases(c)
            cell_bns = _cell_tensors(c)[1]
            for l, b in enumerate(hbs):
                bn = cell_bns[1 + l]
                k = self.bn_modules.index(bn)
                m0 = self.bn_arena.offsets[2 * k]
                v0 = self.bn_arena.offsets[2 * k + 1]
                g0 = self.param_arena.offsets[[id(t) for t in params].index(id(bn.weight))]
                b0 = self.param_arena.offsets[[id(t) for t in params].index(id(bn.bias))]
                n = b.numel()
                self.hidden_biases.append(b)
                hb_mean += list(range(m0, m0 + n)); hb_var += list(range(v0, v0 + n))
                hb_gamma += list(range(g0, g0 + n)); hb_beta += list(range(b0, b0 + n))
            bn_i += len(cell_bns)
        self._hb_idx = None
        if self.hidden_biases:
            self._hb_idx = tuple(torch.tensor(v, dtype=torch.long) for v in (hb_mean, hb_var, hb_gamma, hb_beta))
        self.nbt_arena = _Arena([bn.num_batches_tracked for bn in bns], torch.long)

        desc = _cabi.NisFlowDesc()
        desc.n_flow, desc.n_cells = d, len(cells)
        desc.kind = {"lin": _cabi.KIND_PWLIN, "quad": _cabi.KIND_PWQUAD, "affine": _cabi.KIND_AFFINE}[self.kind]
        desc.n_bins, desc.depth = self.n_bins, len(self.hidden)
        for i, h in enumerate(self.hidden):
            desc.widths[i] = h
        for i, pcol in enumerate(cur):
            desc.out_perm[i] = pcol
        desc.bn_eps, desc.bn_momentum = bns[0].eps, bns[0].momentum
        poff = boff = 0
        for i, (_, c, feed, trafo) in enumerate(cells):
            cd = desc.cells[i]
            cd.n_pass = len(feed)
            for k, col in enumerate(feed):
                cd.feed_idx[k] = col
            for k, col in enumerate(trafo):
                cd.trafo_idx[k] = col
            cd.param_off, cd.bn_off = poff, boff
            p, b = _cell_tensors(c)
            poff += sum(t.numel() for t in p)
            boff += sum(2 * m.num_features for m in b)
        self.desc = desc
        self.n_params = poff
        self._ws = {}
        self._checked = False
        # Set by the training loop around a forward that runs concurrently with other minibatches (manager.py): the running
        # statistics of THAT forward go into this private float32 buffer (bn_arena layout) instead of the shared one, and
        # the loop folds the private buffers back in minibatch order afterwards (BatchNorm's update is a sequential
        # recurrence; concurrent read-modify-writes of the shared buffer would lose updates).
        self.bn_override = None

    # ---- C-ABI plumbing -------------------------------------------------------------------------
    def _self_check(self, lib):
        if self._checked:
            return
        for i in range(self.n_cells):
            p, b = _cell_tensors(self.cells[i][1])
            assert lib.nis_flow_cell_param_count(ctypes.byref(self.desc), i) == sum(t.numel() for t in p)
            assert lib.nis_flow_cell_bn_count(ctypes.byref(self.desc), i) == sum(2 * m.num_features for m in b)
        self._checked = True

    def workspace(self, lib, B, device):
        need = lib.nis_flow_workspace_bytes(ctypes.byref(self.desc), B)
        if need == 0:
            raise _cabi.NisBackendError("flow descriptor rejected by libnisb200")
        key = (device.index, torch.cuda.current_stream(device).cuda_stream)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(int(need * 1.25) + 1024, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    # ---- hidden-layer biases (affine cells) -----------------------------------------------------------
    def _hb(self, device):
        """(index tensors on ``device``, concatenated biases as float32) or None."""
        if not self.hidden_biases:
            return None
        if self._hb_idx[0].device != device:
            self._hb_idx = tuple(t.to(device) for t in self._hb_idx)
        return self._hb_idx, torch.cat([b.detach().reshape(-1).to(device, torch.float32) for b in self.hidden_biases])

    def _bn_for_kernel(self, bn, train, device):
        """The running statistics the kernels should see.  BN(z + b) with running statistics equals BN(z) with the
        running mean lowered by b: eval mode hands the kernel such a copy; train mode (batch statistics: the bias cancels)
        hands it the buffer itself and `_after_train_forward` adds what the bias contributes to the update."""
        hb = self._hb(device)
        if hb is None or train:
            return bn
        eff = bn.clone()
        eff.index_add_(0, hb[0][0], -hb[1])
        return eff

    def _after_train_forward(self, bn, device):
        hb = self._hb(device)
        if hb is not None:                   # running_mean <- (1-m) rm + m (mean(z) + b): the kernel added m mean(z)
            bn.index_add_(0, hb[0][0], float(self.desc.bn_momentum) * hb[1])

    def _hidden_bias_grads(self, gparams, train, device):
        """Gradients of the hidden biases, one tensor per bias.  Train mode: exactly zero (BatchNorm with batch statistics
        removes any constant added in front of it).  Eval mode: y = gamma (z + b - rm) / sqrt(rv + eps) + beta, so
        dL/db = dL/dbeta * gamma / sqrt(rv + eps)."""
        if not self.hidden_biases:
            return []
        if train:
            return [torch.zeros_like(b) for b in self.hidden_biases]
        (_, iv, ig, ib), _ = self._hb(device)
        params = self.param_arena.get(device)
        bn = self.bn_arena.get(device)
        flat = gparams[ib] * params[ig] / torch.sqrt(bn[iv] + float(self.desc.bn_eps))
        out, off = [], 0
        for b in self.hidden_biases:
            out.append(flat[off:off + b.numel()].view(b.shape).to(b.dtype))
            off += b.numel()
        return out

    def bn_saved_count(self, lib):
        return lib.nis_flow_bn_saved_count(ctypes.byref(self.desc))

    def act_saved_count(self, lib, B):
        """Floats of the activation cache a train-mode forward may keep for its backward (0: none).  The cache costs
        cells * depth * width * 4 bytes per point (cfg5: 32 KB); NIS_ACT_CACHE_MAX_BYTES (default 8 GiB) bounds it --
        above the bound the backward recomputes the activations."""
        n = lib.nis_flow_act_saved_count(ctypes.byref(self.desc), B)
        limit = int(os.environ.get("NIS_ACT_CACHE_MAX_BYTES", str(8 << 30)))
        return n if 0 < 4 * n <= limit else 0

    def forward(self, xj, train, want_saved=False, want_bins=False, out_dtype=None):
        """Runs nis_flow_forward(_cached).  Returns (XJ, saved, bn_saved, bins); ``saved`` is a tuple
        (states, activation cache) when the shape keeps one."""
        lib = _cabi.lib()
        self._self_check(lib)
        if xj.dim() != 2 or xj.shape[1] not in (self.n_flow, self.n_flow + 1):
            raise ValueError("expected a [B, %d] or [B, %d] tensor" % (self.n_flow, self.n_flow + 1))
        if not xj.is_cuda:
            dev = self.params[0].device
            if dev.type != "cuda":
                dev = torch.device("cuda", torch.cuda.current_device())
            xj = xj.to(dev)
        dev = xj.device
        xj = xj.detach().contiguous()
        if xj.dtype not in (torch.float32, torch.float64):
            xj = xj.double()
        B, d = xj.shape[0], self.n_flow
        if train and B == 1:
            # batch statistics of one point: torch (and so the reference) refuses, BatchNorm1d in train mode
            raise ValueError("Expected more than 1 value per channel when training, got input size [1, %d]" % d)
        with torch.cuda.device(dev):
            params = self.param_arena.get(dev)
            private = self.bn_override if train else None
            bn_home = private if private is not None else self.bn_arena.get(dev)
            bn = self._bn_for_kernel(bn_home, train, dev)
            out = torch.empty(B, d + 1, dtype=out_dtype or xj.dtype, device=dev)
            saved = torch.empty(self.n_cells + 1, B, d + 1, dtype=torch.float32, device=dev) if want_saved else None
            bn_saved = torch.empty(self.bn_saved_count(lib), dtype=torch.float32, device=dev) \
                if (want_saved and train) else None
            bins = torch.full((self.n_cells, B, d), -1, dtype=torch.int32, device=dev) if want_bins else None
            n_act = self.act_saved_count(lib, B) if (want_saved and train) else 0
            acts = torch.empty(n_act, dtype=torch.float32, device=dev) if n_act else None
            ws = self.workspace(lib, B, dev)
            rc = lib.nis_flow_forward_cached(ctypes.byref(self.desc), _cabi.ptr(params), _cabi.ptr(bn), _cabi.ptr(xj),
                                             _cabi.dtype_code(xj), xj.shape[1], _cabi.ptr(out), _cabi.dtype_code(out),
                                             _cabi.ptr(bins), _cabi.ptr(saved), _cabi.ptr(bn_saved), _cabi.ptr(acts),
                                             _cabi.BN_TRAIN if train else _cabi.BN_EVAL, _cabi.ptr(ws), ws.numel(), B,
                                             _cabi.stream_ptr(dev))
            _cabi.check(rc, "nis_flow_forward_cached")
            if acts is not None:
                saved = (saved, acts)
            if train:
                self._after_train_forward(bn_home, dev)
            if train and private is None:
                self.bn_arena.write_back(bn)
                nbt = self.nbt_arena.get(dev)
                nbt += 1
                self.nbt_arena.write_back(nbt)
        return out, saved, bn_saved, bins

    def inverse(self, yj, train, want_bins=False, out_dtype=None):
        """Runs nis_flow_inverse: yj [B, d(+1)] in the flow's output column order -> (XJ [B, d+1], bins or None), with
        XJ[:, -1] = J_in / prod of the densities (no autograd; train-mode BN uses batch statistics without updating the
        running ones)."""
        lib = _cabi.lib()
        self._self_check(lib)
        if yj.dim() != 2 or yj.shape[1] not in (self.n_flow, self.n_flow + 1):
            raise ValueError("expected a [B, %d] or [B, %d] tensor" % (self.n_flow, self.n_flow + 1))
        if not yj.is_cuda:
            dev = self.params[0].device
            if dev.type != "cuda":
                dev = torch.device("cuda", torch.cuda.current_device())
            yj = yj.to(dev)
        dev = yj.device
        yj = yj.detach().contiguous()
        if yj.dtype not in (torch.float32, torch.float64):
            yj = yj.double()
        B, d = yj.shape[0], self.n_flow
        if train and B == 1:
            raise ValueError("Expected more than 1 value per channel when training, got input size [1, %d]" % d)
        with torch.cuda.device(dev):
            params = self.param_arena.get(dev)
            bn = self._bn_for_kernel(self.bn_arena.get(dev), train, dev)
            out = torch.empty(B, d + 1, dtype=out_dtype or yj.dtype, device=dev)
            bins = torch.full((self.n_cells, B, d), -1, dtype=torch.int32, device=dev) if want_bins else None
            ws = self.workspace(lib, B, dev)
            rc = lib.nis_flow_inverse(ctypes.byref(self.desc), _cabi.ptr(params), _cabi.ptr(bn), _cabi.ptr(yj),
                                      _cabi.dtype_code(yj), yj.shape[1], _cabi.ptr(out), _cabi.dtype_code(out),
                                      _cabi.ptr(bins), _cabi.BN_TRAIN if train else _cabi.BN_EVAL, _cabi.ptr(ws),
                                      ws.numel(), B, _cabi.stream_ptr(dev))
            _cabi.check(rc, "nis_flow_inverse")
        return out, bins

    def backward(self, saved, bn_saved, grad_out, train, need_grad_in):
        """Runs nis_flow_backward(_cached).  Returns (grad_params flat float32, grad_in or None)."""
        lib = _cabi.lib()
        acts = None
        if isinstance(saved, tuple):
            saved, acts = saved
        dev = saved.device
        B, d = saved.shape[1], self.n_flow
        grad_out = grad_out.contiguous()
        if grad_out.dtype not in (torch.float32, torch.float64):
            grad_out = grad_out.double()
        with torch.cuda.device(dev):
            params = self.param_arena.get(dev)
            bn = self._bn_for_kernel(self.bn_arena.get(dev), train, dev)
            gparams = torch.zeros(self.n_params, dtype=torch.float32, device=dev)
            gin = torch.empty(B, d + 1, dtype=grad_out.dtype, device=dev) if need_grad_in else None
            ws = self.workspace(lib, B, dev)
            rc = lib.nis_flow_backward_cached(ctypes.byref(self.desc), _cabi.ptr(params), _cabi.ptr(bn), _cabi.ptr(saved),
                                              _cabi.ptr(bn_saved), _cabi.ptr(acts), _cabi.ptr(grad_out),
                                              _cabi.dtype_code(grad_out), _cabi.ptr(gparams), _cabi.ptr(gin),
                                              _cabi.BN_TRAIN if train else _cabi.BN_EVAL, _cabi.ptr(ws), ws.numel(), B,
                                              _cabi.stream_ptr(dev))
            _cabi.check(rc, "nis_flow_backward_cached")
        return gparams, gin


class _FlowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xj, spec, train, need, *params):
        # ``need`` comes from flow_apply: grad mode is always off inside Function.forward and
        # ctx.needs_input_grad only reflects requires_grad, so under torch.no_grad() (integrate, the tail
        # integration, create_model's trial pass) neither can tell that no backward will follow
        out, saved, bn_saved, _ = spec.forward(xj, train, want_saved=need)
        acts = None
        if isinstance(saved, tuple):
            saved, acts = saved
        ctx.spec, ctx.train = spec, train
        ctx.in_cols, ctx.in_dtype, ctx.in_device = xj.shape[1], xj.dtype, xj.device
        ctx.save_for_backward(*[t for t in (saved, bn_saved, acts) if t is not None])
        ctx.has_bn_saved = bn_saved is not None
        ctx.has_acts = acts is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        spec = ctx.spec
        tensors = ctx.saved_tensors
        saved = tensors[0]
        bn_saved = tensors[1] if ctx.has_bn_saved else None
        if ctx.has_acts:
            saved = (saved, tensors[-1])
        gparams, gin = spec.backward(saved, bn_saved, grad_out, ctx.train, ctx.needs_input_grad[0])
        if gin is not None:
            gin = gin[:, :ctx.in_cols].to(device=ctx.in_device, dtype=ctx.in_dtype)
        grads = []
        for p, off, need in zip(spec.params, spec.param_arena.offsets, ctx.needs_input_grad[4:]):
            grads.append(gparams[off:off + p.numel()].view(p.shape).to(p.dtype) if need else None)
        grads += spec._hidden_bias_grads(gparams, ctx.train, tensors[0].device)
        return (gin, None, None, None) + tuple(grads)


def flow_apply(spec, xj, train):
    """Differentiable fused flow: [B, d(+1)] -> [B, d+1]."""
    if xj.is_cuda:
        # The parameters follow the input's device (they are re-homed as views of one arena there).  Do it
        # before autograd records them as inputs, or the first backward after a device change sees gradients
        # on another device than the one it noted for the parameters.
        spec.param_arena.get(xj.device)
    need = torch.is_grad_enabled() and (xj.requires_grad or any(p.requires_grad for p in spec.params)
                                        or any(p.requires_grad for p in spec.hidden_biases))
    return _FlowFn.apply(xj, spec, bool(train), need, *spec.params, *spec.hidden_biases)


class FlowSequential(torch.nn.Sequential):
    """What ``create_model`` returns as ``_model``: a ``torch.nn.Sequential`` whose children carry the
    reference's names ('0', 'roll0', 'mask0', ...), so ``parameters()`` / ``state_dict()`` /
    ``deepcopy`` behave as in the reference, but whose ``forward`` is ONE fused C-ABI call."""

    def __init__(self, n_flow):
        super().__init__()
        self.n_flow = n_flow
        self._spec = None
        self._spec_key = None

    def spec(self):
        key = tuple((n, id(m)) for n, m in self.named_children())
        if self._spec is None or self._spec_key != key:
            self._spec = FlowSpec(list(self.named_children()), self.n_flow)
            self._spec_key = key
        return self._spec

    def forward(self, input):
        return flow_apply(self.spec(), input, self.training)

    def inverse(self, input):
        """The inverse map (a to-do in the reference, README.md:68-69): [B, d(+1)] points in the flow's
        output space -> [B, d+1] latent points with the Jacobian column divided by the product of the densities, so that
        ``model.inverse(model(x))`` returns x with Jacobian 1.  Not differentiable."""
        with torch.no_grad():
            return self.spec().inverse(input, self.training)[0]

    def forward_with_bins(self, input):
        """(XJ, bins[n_cells, B, n_flow] int32) without autograd — for parity checks."""
        out, _, _, bins = self.spec().forward(input, self.training, want_bins=True)
        return out, bins

    def __deepcopy__(self, memo):
        new = torch.nn.Sequential.__new__(type(self))
        memo[id(self)] = new
        import copy
        for k, v in self.__dict__.items():
            if k in ("_spec", "_spec_key"):
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        new._spec = None
        new._spec_key = None
        return new
