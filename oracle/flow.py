"""Oracle: coupling-cell normalizing flow (PWLin / PWQuad) — CPU, float64, closed form.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Differentiable torch code so that
``torch.autograd`` of this restatement is the gradient oracle for the fused CUDA backward.

Reference files restated here (all paths relative to /root/reference/nisrep/normalizing_flows):
  layers/coupling_cells.py:73-142   PWLin
  layers/coupling_cells.py:144-228  PWQuad
  layers/coupling_cells.py:230-254  RectNN conditioner (PWLin inlines the same stack, :84-104)
  layers/layers.py:6-51,66-91       MaskLayer / DeMaskLayer / AddJacobian / RollLayer
  manager.py:20-36                  get_bin
  manager.py:474-499                PWLinManager.create_model   (topology only)
  manager.py:518-600                PWQuadManager.create_model  (topology only)

State layout is the reference's: rows are points, columns ``0..d-1`` the coordinates, column ``d``
the accumulated Jacobian (a product, not a log).
"""
import math

import torch

BN_EPS = 1e-5        # torch.nn.BatchNorm1d default, used by coupling_cells.py:236-246
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------
# topology (host logic of create_model)
# ----------------------------------------------------------------------------------------------
def get_bin(x, n=0):
    """manager.py:20-36 — binary digits of x, MSB first, zero-filled to n digits."""
    return [int(c) for c in format(int(x), "b").zfill(n)]


def pwlin_layers(n_flow, n_pass_through, n_cells, roll_step):
    """manager.py:487-492.  Every roll is registered under the same module name "roll", so
    ``add_module`` replaces it in place: exactly one roll survives, right after cell 0."""
    layers = []
    for i in range(n_cells):
        layers.append(dict(type="cell", name=str(i), P=n_pass_through))
        if i == 0:
            layers.append(dict(type="roll", name="roll", shift=roll_step))
    return layers


def pwquad_n_cells(n_flow, n_cells):
    """manager.py:526-534 — cell-count fix-up.  Returns (n_cells, adjusted?)."""
    if n_cells < 2 * math.ceil(math.log2(n_flow)) and n_cells < n_flow:
        if n_flow <= 6:
            n = n_flow
        elif n_flow == 7:
            n = 6
        else:
            n = int(2 * math.ceil(math.log2(n_flow)))
        return n, True
    return n_cells, False


def pwquad_layers(n_flow, n_cells):
    """manager.py:538-585 — roll layout for d<=7, binary-mask layout (+ extra roll cells) for d>=8."""
    n_cells, _ = pwquad_n_cells(n_flow, n_cells)
    d = n_flow
    layers = []
    if d <= 7:
        P = 1 if d <= 6 else 2
        for i in range(n_cells):
            layers.append(dict(type="cell", name=str(i), P=P))
            if i < n_cells - 1:
                layers.append(dict(type="roll", name="roll%d" % i, shift=1))
            else:
                layers.append(dict(type="roll", name="roll%d" % i, shift=d - ((n_cells - 1) % d)))
    else:
        P = d // 2
        nbits = len(get_bin(d - 1, 0))
        dims_bin = [get_bin(i, nbits) for i in range(d)]
        for c in range(2 * nbits):
            feed, pos = c % 2, c // 2                      # layers.py:15-20
            feeder = [i for i in range(d) if dims_bin[i][pos] == feed]
            trafoer = [i for i in range(d) if dims_bin[i][pos] != feed]
            layers.append(dict(type="mask", name="mask%d" % c, feeder=feeder, trafoer=trafoer))
            layers.append(dict(type="cell", name=str(c), P=len(feeder)))
            layers.append(dict(type="demask", name="demask%d" % c, feeder=feeder, trafoer=trafoer))
        extra = n_cells - 2 * nbits
        for i in range(extra):
            c = i + 2 * nbits
            layers.append(dict(type="cell", name=str(c), P=P))
            if i < extra - 1:
                layers.append(dict(type="roll", name="roll%d" % c, shift=1))
            else:
                layers.append(dict(type="roll", name="roll%d" % c, shift=d - ((extra - 1) % d)))
    return layers


def compile_layers(layers, d):
    """Fold Roll / Mask / DeMask into per-cell column tables over a state that never moves.

    Returns (cells, out_perm): cell c conditions on physical columns ``feed_idx`` and transforms
    physical columns ``trafo_idx`` (in the order the reference cell sees them); the reference's
    final logical column i is physical column ``out_perm[i]``.
    """
    cur = list(range(d))            # cur[logical position] = physical column
    stack = []
    cells = []
    for L in layers:
        t = L["type"]
        if t == "cell":
            P = L["P"]
            cells.append(dict(name=L["name"], P=P, feed_idx=cur[:P], trafo_idx=cur[P:]))
        elif t == "roll":           # layers.py:90-91: out[:, (i+shift)%d] = x[:, i]
            s = L["shift"]
            new = [None] * d
            for i in range(d):
                new[(i + s) % d] = cur[i]
            cur = new
        elif t == "mask":           # layers.py:27-32
            stack.append(cur)
            cur = [cur[i] for i in L["feeder"]] + [cur[i] for i in L["trafoer"]]
        elif t == "demask":         # layers.py:43-51
            cur = stack.pop()
        else:
            raise ValueError(t)
    return cells, cur


# ----------------------------------------------------------------------------------------------
# conditioner
# ----------------------------------------------------------------------------------------------
def _bn(x, sd, key, train, stats):
    """torch.nn.BatchNorm1d forward.  train: batch mean / biased variance; running statistics are
    updated with momentum 0.1 and the unbiased variance (written into ``stats`` if given)."""
    g, b = sd[key + ".weight"], sd[key + ".bias"]
    if train:
        n = x.shape[0]
        mean = x.mean(0)
        var = ((x - mean) ** 2).mean(0)
        if stats is not None:
            rm, rv = sd[key + ".running_mean"], sd[key + ".running_var"]
            stats[key + ".running_mean"] = ((1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean).detach()
            stats[key + ".running_var"] = ((1 - BN_MOMENTUM) * rv
                                          + BN_MOMENTUM * var * n / max(n - 1, 1)).detach()
            stats[key + ".batch_mean"] = mean.detach()
            stats[key + ".batch_var"] = var.detach()
    else:
        mean, var = sd[key + ".running_mean"], sd[key + ".running_var"]
    return (x - mean) / torch.sqrt(var + BN_EPS) * g + b


def n_hidden(sd, cell):
    """Number of hidden layers of cell ``cell``: Linear modules sit at Sequential index 1,4,7,..."""
    depth = 0
    while "%s.NN.%d.weight" % (cell, 1 + 3 * (depth + 1)) in sd:
        depth += 1
    return depth


def rectnn(sd, cell, xA, train, stats=None):
    """coupling_cells.py:230-254: BN -> [Linear(no bias) -> BN -> ReLU]*depth -> Linear(bias)."""
    depth = n_hidden(sd, cell)
    pre = "%s.NN." % cell
    h = _bn(xA, sd, pre + "0", train, stats)
    for l in range(depth):
        h = h @ sd[pre + "%d.weight" % (1 + 3 * l)].T
        if pre + "%d.bias" % (1 + 3 * l) in sd:         # AffineCoupling: torch-default Linear (coupling_cells.py:27-38)
            h = h + sd[pre + "%d.bias" % (1 + 3 * l)]
        h = torch.relu(_bn(h, sd, pre + "%d" % (2 + 3 * l), train, stats))
    last = pre + "%d" % (1 + 3 * depth)
    return h @ sd[last + ".weight"].T + sd[last + ".bias"]


# ----------------------------------------------------------------------------------------------
# coupling cells
# ----------------------------------------------------------------------------------------------
def pwlin_cell(sd, cell, x, P, n_bins, train, stats=None, edges=None, clamp_bins=False):
    """coupling_cells.py:107-142.  x: [B, d+1].  Returns (out [B, d+1], bins [B, T]).
    ``edges`` (list) receives the distance [B, T] of every transformed coordinate to its nearest bin edge.
    ``clamp_bins``: a coordinate of exactly 1.0 (the output of a previous cell can round to it) makes the reference
    gather out of range and raise (:126-133, no clamp); with this flag the bin is clamped to n_bins-1 like the
    CUDA kernels do (known divergences) so that large batches can be compared point by point."""
    d = x.shape[1] - 1
    T = d - P
    xA, xB, J = x[:, :P], x[:, P:d], x[:, d]
    Z = rectnn(sd, cell, xA, train, stats).reshape(-1, T, n_bins)
    Q = torch.exp(Z)                                    # :115 (no max subtraction)
    Qsum = torch.cumsum(Q, -1)
    norm = Qsum[:, :, -1:]
    Q = Q / (norm / n_bins)                             # :121 bin heights (pdf)
    C = torch.cat((torch.zeros_like(norm), Qsum / norm), -1)   # :123-124 cdf at left edges
    a = xB * n_bins
    bins = torch.floor(a)
    if clamp_bins:
        bins = bins.clamp(0, n_bins - 1)
    if edges is not None:
        edges.append(((a - torch.round(a)).abs() / n_bins).detach())
    alpha = (a - bins) / n_bins                         # :130-131
    k = bins.long().unsqueeze(-1)
    Qk = torch.gather(Q, -1, k).squeeze(-1)
    Ck = torch.gather(C, -1, k).squeeze(-1)
    y = Qk * alpha + Ck                                 # :139
    J = J * torch.prod(Qk, -1)                          # :141
    return torch.cat((xA, y, J.unsqueeze(-1)), -1), bins.long()


def pwquad_cell(sd, cell, x, P, n_bins, train, stats=None, edges=None, cond=None):
    """coupling_cells.py:159-228.  Returns (out [B, d+1], bins [B, T]).
    ``edges`` (list) receives the distance [B, T] of every transformed coordinate to its nearest bin edge; ``cond``
    (list) the sensitivity [B] of this cell's log-Jacobian to its input coordinates, sum_t |d log f_t / d x_t| =
    |V_{k+1} - V_k| / (W_k lerp(V_k, V_{k+1}, alpha)): what a rounding error of x is multiplied by inside a narrow bin."""
    d = x.shape[1] - 1
    T = d - P
    nb = n_bins
    xA, xB, J = x[:, :P], x[:, P:d], x[:, d]
    xB = torch.where(xB > 1 - 1e-6, torch.full_like(xB, 1 - 1e-6), xB)     # :167
    Z = rectnn(sd, cell, xA, train, stats).reshape(-1, T, 2 * nb + 1)
    V = torch.exp(Z[:, :, :nb + 1])                     # :173,:189 vertex heights
    W = torch.exp(Z[:, :, nb + 1:])                     # :175,:178 bin widths
    Wsum = torch.cumsum(W, -1)
    Wn = Wsum[:, :, -1:]
    W = W / Wn                                          # :184
    Wsum = Wsum / Wn                                    # :186 right edges E_1..E_nb
    area = torch.cumsum((V[:, :, :-1] + V[:, :, 1:]) / 2 * W, -1)           # :194
    V = V / area[:, :, -1:]                             # :196-197
    E = torch.cat((torch.zeros_like(Wn), Wsum), -1)     # :198 left edges E_0..E_nb
    # :199-202 — argmax(cat(1e-30, (Wsum<=xB)*Wsum)) == number of right edges <= xB
    k = (Wsum <= xB.unsqueeze(-1)).sum(-1, keepdim=True)
    if edges is not None:
        edges.append((E - xB.unsqueeze(-1)).abs().min(-1).values.detach())
    Wk = torch.gather(W, -1, k).squeeze(-1)
    alpha = (xB - torch.gather(E, -1, k).squeeze(-1)) / Wk                  # :206-207
    S = torch.cat((torch.zeros_like(Wn),
                   torch.cumsum((V[:, :, :-1] + V[:, :, 1:]) / 2 * W, -1)), -1)   # :209-210
    Vk = torch.gather(V, -1, k).squeeze(-1)
    Vk1 = torch.gather(V, -1, k + 1).squeeze(-1)
    y = alpha ** 2 / 2 * ((Vk1 - Vk) * Wk) + alpha * Vk * Wk + torch.gather(S, -1, k).squeeze(-1)
    if cond is not None:
        cond.append(((Vk1 - Vk).abs() / (Wk * torch.lerp(Vk, Vk1, alpha))).sum(-1).detach())
    J = J * torch.prod(torch.lerp(Vk, Vk1, alpha), -1)  # :224-225
    return torch.cat((xA, y, J.unsqueeze(-1)), -1), k.squeeze(-1)


def affine_cell(sd, cell, x, P, n_bins, train, stats=None, edges=None):
    """AffineCoupling.forward, coupling_cells.py:49-70.  Returns (out [B, d+1], bins [B, T] of zeros).
    Z = NN(xA) reshaped (2, T); s0 = exp(Z[:, 0]); s1 = relu(Z[:, 1]); yB = atan(20 s0 xB + s1) / (pi/2);
    J *= prod(20 s0) * (1 / (pi/2)) * prod(1 / (v^2 + 1)) - the 1/(pi/2) ONCE per cell, as the reference writes it."""
    d = x.shape[1] - 1
    T = d - P
    xA, xB, J = x[:, :P], x[:, P:d], x[:, d]
    Z = rectnn(sd, cell, xA, train, stats).reshape(-1, 2, T)
    s0 = torch.exp(Z[:, 0])
    s1 = torch.where(Z[:, 1] > 0, Z[:, 1], torch.zeros_like(Z[:, 1]))
    v = xB * (20 * s0) + s1
    diff = 1 / (v ** 2 + 1)
    y = torch.atan(v) / (math.pi / 2)
    J = J * torch.prod(20 * s0, 1) * (1 / (math.pi / 2)) * torch.prod(diff, 1)
    return torch.cat((xA, y, J.unsqueeze(-1)), -1), torch.zeros(x.shape[0], T, dtype=torch.long)


def affine_cell_inverse(sd, cell, yj, P, n_bins, train, stats=None):
    """Inverse of affine_cell: yj = (xA, y, J) -> (xA, x, J / (cell Jacobian))."""
    d = yj.shape[1] - 1
    T = d - P
    xA, yB, J = yj[:, :P], yj[:, P:d], yj[:, d]
    Z = rectnn(sd, cell, xA, train, stats).reshape(-1, 2, T)
    s0 = torch.exp(Z[:, 0])
    s1 = torch.relu(Z[:, 1])
    v = torch.tan(yB * (math.pi / 2))
    x = (v - s1) / (20 * s0)
    J = J / (torch.prod(20 * s0, 1) * (1 / (math.pi / 2)) * torch.prod(1 / (v ** 2 + 1), 1))
    return torch.cat((xA, x, J.unsqueeze(-1)), -1), torch.zeros(yj.shape[0], T, dtype=torch.long)


# ----------------------------------------------------------------------------------------------
# inverse cells.  The reference has no inverse (its README lists it as to do, README.md:68-69); these are
# the algebraic inverses of the two maps above, validated by round trips against the reference-pinned forward.
# ----------------------------------------------------------------------------------------------
def pwlin_cell_inverse(sd, cell, yj, P, n_bins, train, stats=None):
    """Inverse of pwlin_cell: yj = (xA, y, J) -> (xA, x, J / prod Q_k).  The conditioner sees the same xA."""
    d = yj.shape[1] - 1
    T = d - P
    xA, yB, J = yj[:, :P], yj[:, P:d], yj[:, d]
    Z = rectnn(sd, cell, xA, train, stats).reshape(-1, T, n_bins)
    Q = torch.exp(Z)
    Qsum = torch.cumsum(Q, -1)
    norm = Qsum[:, :, -1:]
    Q = Q / (norm / n_bins)
    C = torch.cat((torch.zeros_like(norm), Qsum / norm), -1)         # cdf at left edges, C[..., n_bins] = 1
    k = ((C[:, :, 1:-1] <= yB.unsqueeze(-1)).sum(-1)).clamp(0, n_bins - 1).unsqueeze(-1)   # largest k with C_k <= y
    Qk = torch.gather(Q, -1, k).squeeze(-1)
    Ck = torch.gather(C, -1, k).squeeze(-1)
    x = k.squeeze(-1).to(yB.dtype) / n_bins + (yB - Ck) / Qk
    J = J / torch.prod(Qk, -1)
    return torch.cat((xA, x, J.unsqueeze(-1)), -1), k.squeeze(-1)


def pwquad_cell_inverse(sd, cell, yj, P, n_bins, train, stats=None):
    """Inverse of pwquad_cell (away from the forward's clamp at 1 - 1e-6): solve the bin's quadratic for alpha."""
    d = yj.shape[1] - 1
    T = d - P
    nb = n_bins
    xA, yB, J = yj[:, :P], yj[:, P:d], yj[:, d]
    Z = rectnn(sd, cell, xA, train, stats).reshape(-1, T, 2 * nb + 1)
    V = torch.exp(Z[:, :, :nb + 1])
    W = torch.exp(Z[:, :, nb + 1:])
    Wsum = torch.cumsum(W, -1)
    Wn = Wsum[:, :, -1:]
    W = W / Wn
    Wsum = Wsum / Wn
    area = torch.cumsum((V[:, :, :-1] + V[:, :, 1:]) / 2 * W, -1)
    V = V / area[:, :, -1:]
    E = torch.cat((torch.zeros_like(Wn), Wsum), -1)
    S = torch.cat((torch.zeros_like(Wn), torch.cumsum((V[:, :, :-1] + V[:, :, 1:]) / 2 * W, -1)), -1)   # cdf at the edges
    k = ((S[:, :, 1:-1] <= yB.unsqueeze(-1)).sum(-1)).clamp(0, nb - 1).unsqueeze(-1)
    Wk = torch.gather(W, -1, k).squeeze(-1)
    Vk = torch.gather(V, -1, k).squeeze(-1)
    Vk1 = torch.gather(V, -1, k + 1).squeeze(-1)
    c = yB - torch.gather(S, -1, k).squeeze(-1)
    a = (Vk1 - Vk) * Wk / 2
    b = Vk * Wk
    alpha = 2 * c / (b + torch.sqrt(b * b + 4 * a * c))
    x = torch.gather(E, -1, k).squeeze(-1) + alpha * Wk
    J = J / torch.prod(torch.lerp(Vk, Vk1, alpha), -1)
    return torch.cat((xA, x, J.unsqueeze(-1)), -1), k.squeeze(-1)


def flow_inverse(layers, sd, yj, kind, n_bins, train=False, stats=None):
    """Inverse of flow_forward: the Sequential backwards (inverse rolls / masks, inverse cells).  yj: [B, d+1] in
    the reference's output column order; returns (XJ, bins) with XJ[:, -1] = J_in / prod f, so that
    flow_inverse(flow_forward(x)) = x with Jacobian 1."""
    cell_fn = {"lin": pwlin_cell_inverse, "quad": pwquad_cell_inverse, "affine": affine_cell_inverse}[kind]
    x = yj
    all_bins = []
    for L in reversed(layers):
        t = L["type"]
        if t == "cell":
            x, b = cell_fn(sd, L["name"], x, L["P"], n_bins, train, stats)
            all_bins.append(b)
        elif t == "roll":
            x = torch.cat((torch.roll(x[:, :-1], -L["shift"], -1), x[:, -1:]), -1)
        elif t == "demask":                 # forward demask scattered (feeder, trafoer) back: undo = mask
            x = torch.cat((x[:, L["feeder"]], x[:, L["trafoer"]], x[:, -1:]), -1)
        elif t == "mask":                   # forward mask gathered: undo = scatter back
            ret = torch.empty_like(x[:, :-1])
            ret[:, L["feeder"] + L["trafoer"]] = x[:, :-1]
            x = torch.cat((ret, x[:, -1:]), -1)
    return x, all_bins[::-1]


# ----------------------------------------------------------------------------------------------
# whole flow, layer by layer like the reference Sequential
# ----------------------------------------------------------------------------------------------
def flow_forward(layers, sd, xj, kind, n_bins, train=True, stats=None, trace=None, edges=None, clamp_bins=False, cond=None):
    """Run the reference Sequential.  xj: [B, d+1] float64.  Returns (XJ [B, d+1], bins [B, C, T_c]
    as a list of per-cell LongTensors).  ``trace`` (dict) receives each module's output by name, ``edges`` (list)
    each cell's [B, T_c] distances to the nearest bin edge, ``cond`` (list, PWQuad) each cell's [B] sensitivity of the
    log-Jacobian to its input coordinates (see pwquad_cell)."""
    d = xj.shape[1] - 1
    cell_fn = {"lin": pwlin_cell, "quad": pwquad_cell, "affine": affine_cell}[kind]
    x = xj
    all_bins = []
    for L in layers:
        t = L["type"]
        if t == "cell":
            extra = {"clamp_bins": True} if clamp_bins and kind == "lin" else {}
            if cond is not None and kind == "quad":
                extra["cond"] = cond
            x, b = cell_fn(sd, L["name"], x, L["P"], n_bins, train, stats, edges, **extra)
            all_bins.append(b)
        elif t == "roll":
            x = torch.cat((torch.roll(x[:, :-1], L["shift"], -1), x[:, -1:]), -1)
        elif t == "mask":
            x = torch.cat((x[:, L["feeder"]], x[:, L["trafoer"]], x[:, -1:]), -1)
        elif t == "demask":
            ret = torch.empty_like(x[:, :-1])
            ret[:, L["feeder"] + L["trafoer"]] = x[:, :-1]
            x = torch.cat((ret, x[:, -1:]), -1)
        if trace is not None:
            trace[L["name"]] = x.detach().clone()
    return x, all_bins


def flow_forward_compiled(cells, out_perm, sd, xj, kind, n_bins, train=True, stats=None):
    """Same map, but in the index-table form the CUDA kernels use (state never moves; each cell
    reads/writes physical columns).  Used to check that folding Roll/Mask/DeMask is exact."""
    d = xj.shape[1] - 1
    cell_fn = {"lin": pwlin_cell, "quad": pwquad_cell, "affine": affine_cell}[kind]
    state = xj.clone()
    all_bins = []
    for c in cells:
        cols = c["feed_idx"] + c["trafo_idx"] + [d]
        out, b = cell_fn(sd, c["name"], state[:, cols], c["P"], n_bins, train, stats)
        new = state.clone()
        new[:, cols] = out
        state = new
        all_bins.append(b)
    return state[:, out_perm + [d]], all_bins


# ----------------------------------------------------------------------------------------------
# parameter initialisation identical in distribution to torch.nn defaults (for full-size cases
# whose weights are too large to commit as fixtures: both sides load the SAME generated dict)
# ----------------------------------------------------------------------------------------------
def init_state_dict(cells, d, kind, n_bins, NN, seed, dtype=torch.float64, bn_jitter=0.0):
    """kaiming-uniform(a=sqrt(5)) Linear weights (bound 1/sqrt(fan_in)), BN gamma=1, beta=0,
    running_mean=0, running_var=1 — torch.nn.Linear / BatchNorm1d defaults.  ``bn_jitter`` > 0
    perturbs the BN affine and running statistics so eval-mode parity is not trivially satisfied."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def U(shape, bound):
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    def bn(key, n):
        sd[key + ".weight"] = torch.ones(n, dtype=dtype) + bn_jitter * U((n,), 1.0)
        sd[key + ".bias"] = bn_jitter * U((n,), 1.0)
        sd[key + ".running_mean"] = bn_jitter * U((n,), 1.0)
        sd[key + ".running_var"] = torch.ones(n, dtype=dtype) + bn_jitter * U((n,), 0.5)
        sd[key + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    for c in cells:
        P = c["P"]
        T = d - P
        out = T * {"lin": n_bins, "quad": 2 * n_bins + 1, "affine": 2}[kind]
        pre = "%s.NN." % c["name"]
        bn(pre + "0", P)
        fan = P
        for l, h in enumerate(NN):
            sd[pre + "%d.weight" % (1 + 3 * l)] = U((h, fan), 1 / math.sqrt(fan))
            if kind == "affine":
                sd[pre + "%d.bias" % (1 + 3 * l)] = U((h,), 1 / math.sqrt(fan))
            bn(pre + "%d" % (2 + 3 * l), h)
            fan = h
        sd[pre + "%d.weight" % (1 + 3 * len(NN))] = U((out, fan), 1 / math.sqrt(fan))
        sd[pre + "%d.bias" % (1 + 3 * len(NN))] = U((out,), 1 / math.sqrt(fan))
    return sd
