"""Model factories, variance-loss training and Monte Carlo integration — the public surface of
nisrep/normalizing_flows/manager.py (ModelAPI, BasicManager, PWLinManager, PWQuadManager) on top of the
fused sm_100a kernels.

Same method names, positional argument order, attributes (``_model``, ``best_model``, ``format_input``,
``best_loss``, ``best_loss_rel``, ``best_func_count``, ``varJ``, ``DKL``, ``best_var``, ``best_epoch``,
``int_loss``, ``history``, ``integ_tot``, ``err_tot``) and return conventions as the reference
(manager.py:66-70, 380, 474-480, 518-523).  What differs, on purpose:

* the flow forward/backward is one fused C-ABI call per (mini)batch instead of ~40 ATen launches per cell;
* no ``gc.collect()`` per minibatch (manager.py:270 — 90 % of the reference's wall time, no effect on results);
* when ``torch.distributed`` is initialised the minibatches of an epoch are dealt round-robin to the
  ranks, gradients and the epoch loss are sum-allreduced (the reference objective is a mean over
  minibatches with per-minibatch BN statistics, so this is the same objective), ``maxf`` is
  max-allreduced and ``integrate`` sum-allreduces (sum w, sum w^2, n) per iteration;
* parameters are float32 (the kernels compute in fp32); tensors handed to the user integrand ``f`` are
  float64 like the reference's.
"""
import copy
import datetime
import math
import os
import warnings
import weakref

import numpy as np
import torch
import torch.distributed as dist
from tqdm.autonotebook import tqdm

from .. import _cabi
from ..flowspec import FlowSequential
from .layers.coupling_cells import AffineCoupling, PWLin, PWQuad
from .layers.layers import AddJacobian, DeMaskLayer, MaskLayer, RollLayer
from .misc import tqdm_recycled


class NonFiniteWeightWarning(RuntimeWarning):
    """Some f(x)*J(x) handed to ``integrate`` / ``weight_statistics`` were inf or NaN."""


def get_bin(x, n=0):
    """Binary digits of x, most significant first, left-padded with zeros to n digits (manager.py:20-36)."""
    return [int(ch) for ch in format(x, "b").zfill(n)]


def normal(x, mu, sigma, n_flow):
    """Isotropic Gaussian density helper (manager.py:39-40)."""
    return torch.exp(-torch.sum((x - mu) ** 2 / (2 * sigma ** 2), -1)) / (sigma * np.sqrt((2 * np.pi) ** n_flow))


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world):
    """[first, first+count) of rank's share when n Monte Carlo points are dealt to `world` ranks."""
    count = n // world + (1 if rank < n % world else 0)
    first = (n // world) * rank + min(rank, n % world)
    return first, count


def _device(dev):
    if not torch.cuda.is_available():
        raise _cabi.NisBackendError("nf_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if isinstance(dev, torch.device):
        return dev
    return torch.device("cuda:" + str(0 if dev is None else dev))


class ModelAPI():

    @property
    def model(self):
        if self._model is not None:
            return self._model
        raise AttributeError("No model was instantiated")


class EpochState:
    """Best-model bookkeeping, preburn switch and early stopping of the training loop
    (manager.py:205-210, 291-327) as a pure host state machine over the epoch losses."""

    def __init__(self, int_loss, preburn_time, kill_counter, impr_ratio):
        self.check_time = preburn_time if preburn_time > 10 else 50
        self.preburn_time, self.kill_counter, self.impr_ratio = preburn_time, kill_counter, impr_ratio
        self.int_loss = int_loss
        self.best_loss = int_loss
        self.stale_save = 1000
        self.preburner = preburn_time > 0
        self.counter = 0
        self.last_loss = 1000

    def improved(self, loss, track):
        """True when this epoch's model must be snapshotted as best_model (manager.py:293)."""
        if track and loss < self.best_loss and not self.preburner:
            self.best_loss = loss
            return True
        return False

    def advance(self, i, loss):
        """Update counters after epoch i; True = stop training (manager.py:307-327)."""
        if loss < self.last_loss:
            self.counter = 0
        else:
            self.counter += 1
            if self.counter > self.kill_counter:
                if not self.preburner:
                    return True
                self.counter = 0
                self.preburner = False
        self.last_loss = loss
        on_check = i % self.check_time == 0
        if on_check and i > self.preburn_time + 1 and not self.preburner and \
                float(self.best_loss / self.stale_save) > 1 - self.impr_ratio:
            return True
        if on_check and not self.preburner and (self.best_loss < self.int_loss or i > 300):
            self.stale_save = self.best_loss
        if self.preburner and (loss < 0.25 * self.best_loss or i > self.preburn_time):
            self.preburner = False
        return False


class BasicManager(ModelAPI):
    """Training (variance loss, Jacobian from the forward pass) and integration."""

    format_input = AddJacobian()

    def __init__(self, n_flow=2, *args):
        self.n_flow = n_flow
        self._model = None
        self._inverse_model = None
        self.optimizer_object = None
        self.best_model = None

    # ------------------------------------------------------------------------------------------
    def _uniform(self, n, dev, dtype=torch.double, generator=None):
        """Latent points: torch's device generator, so ``torch.manual_seed`` controls the stream; under
        torch.distributed the per-rank generator of ``_rank_generator``."""
        return torch.rand(n, self.n_flow, device=dev, dtype=dtype, generator=generator)

    def _snapshot_best(self):
        """``best_model = deepcopy(model)`` (manager.py:299) without rebuilding the module tree on every improving
        epoch (a deepcopy of the Sequential costs ~6 ms: most of the README example's wall time once the kernels are
        fast).  After the first deepcopy the snapshot is refreshed by copying the three flat arenas (parameters, BN
        running statistics, batch counters) - three small device copies."""
        src, dst = self.model, self.best_model
        fast = isinstance(src, FlowSequential) and isinstance(dst, FlowSequential) and dst is not src and \
            getattr(self, "_best_src", None) is not None and self._best_src() is src and dst.training == src.training
        if fast:
            try:
                ss, ds = src.spec(), dst.spec()
                dev = ss.params[0].device
                pairs = []
                for a, b in ((ss.param_arena, ds.param_arena), (ss.bn_arena, ds.bn_arena), (ss.nbt_arena, ds.nbt_arena)):
                    fa, fb = a.get(dev), b.get(dev)
                    if a.flat is None or b.flat is None or fa.shape != fb.shape:
                        raise RuntimeError("arena not linked")
                    pairs.append((fb, fa))
                with torch.no_grad():
                    for fb, fa in pairs:
                        fb.copy_(fa)
                    for hb, ha in zip(ds.hidden_biases, ss.hidden_biases):      # affine cells: biases outside the arenas
                        hb.copy_(ha)
                return
            except Exception:
                pass
        self.best_model = copy.deepcopy(src)
        self._best_src = weakref.ref(src)

    @staticmethod
    def _rank_generator(dev, rank, world):
        """Ranks seeded alike (the usual ``torch.manual_seed(s)`` on every rank) would all draw the same latent
        points and the summed gradient would be one minibatch repeated ``world`` times.  So under
        torch.distributed the latent points come from a per-rank device generator seeded with (a draw from rank
        0's CPU generator) + rank: reproducible from the user's seed, different on every rank."""
        if world == 1:
            return None
        base = torch.empty((), dtype=torch.int64).random_(0, 2 ** 62).to(dev)
        dist.broadcast(base, 0)
        g = torch.Generator(device=dev)
        g.manual_seed(int(base.item()) + rank)
        return g

    def _train_variance_forward_seq(self, f, optimizer_object, log=True, logdir=None, batch_size=10000, epochs=10,
                                    epoch_start=0, pretty_progressbar=True, save_best=True, run=None, dev=0,
                                    mini_batch_size=2000, integrate=False, preburn_time=75, kill_counter=7,
                                    impr_ratio=1e-2, loss_mode="var"):
        """Train on the variance of f(x) J(x) / maxf over fresh uniform minibatches (manager.py:66-378)."""
        dev = _device(dev)
        rank, world = _world()
        if mini_batch_size > batch_size:
            mini_batch_size = batch_size
        n_minibatches = int(batch_size / mini_batch_size)
        batch_size = batch_size - (batch_size % mini_batch_size)
        my_minibatches = [j for j in range(n_minibatches) if j % world == rank]

        filename = None
        if log and rank == 0:
            base = os.path.join(logdir, str(run._id)) if run is not None else logdir
            filename = os.path.join(base, "torch")
            try:
                os.makedirs(base, exist_ok=True)
                torch.save({'model_state_dict': self.best_model.state_dict()}, filename + "_int")
            except Exception:
                print("Torch save not possible")

        integ = torch.zeros((epochs + 1,), device=dev)
        err = torch.zeros((epochs + 1,), device=dev)
        if pretty_progressbar and rank == 0:
            epoch_progress = tqdm(range(epoch_start, epoch_start + epochs), leave=False,
                                  desc="Loss: {0:.3e} | Epoch".format(0.))
            minibatch_progress = tqdm_recycled(my_minibatches, leave=False, desc="Step") \
                if len(my_minibatches) > 1 else my_minibatches
        else:
            epoch_progress = range(epoch_start, epoch_start + epochs)
            minibatch_progress = my_minibatches

        self.model.to(dev)
        if world > 1:
            self._sync_model()
        if loss_mode not in ("var", "est"):
            print("Unknown loss function")
            return

        # ---- initial loss of the untransformed integrand, maxf (manager.py:139-165) ---------------
        self.best_loss = 0
        self.best_var = 0
        maxf = torch.zeros((), device=dev, dtype=torch.double)
        w = None
        gen = self._rank_generator(dev, rank, world)
        for _ in range(self.n_flow):
            w = self._uniform(2 * mini_batch_size, dev, generator=gen)
            fres = f(w)
            integ[0] += torch.sum(fres) / (self.n_flow * 2 * mini_batch_size)
            err[0] += torch.var(fres) / self.n_flow
            maxf = torch.maximum(maxf, torch.max(fres).to(maxf.dtype))
            if world > 1:
                dist.all_reduce(maxf, op=dist.ReduceOp.MAX)
            if loss_mode == "var":
                self.best_loss = self.best_loss + torch.var(fres / maxf).detach() / self.n_flow
            else:
                self.best_loss = self.best_loss + torch.mean(fres ** 2).detach() / self.n_flow
            self.best_var += float((torch.var((fres / maxf) ** 2) / 2 * mini_batch_size).detach())

        if world > 1:                       # every rank must take the same early-stopping decisions
            init = torch.stack((torch.as_tensor(self.best_loss, device=dev).double(), integ[0].double(), err[0].double()))
            dist.all_reduce(init)
            init /= world
            self.best_loss, integ[0], err[0] = init[0], init[1], init[2]
        if save_best or log:
            XJ = self.model(self.format_input(w, dev))
            X, J = XJ[:, :-1], XJ[:, -1]
            self.varJ = torch.mean(J ** 2).detach()
            self.DKL = torch.nn.KLDivLoss(reduction='batchmean')(torch.log(X + 1e-45), w).detach()
            self._best_src = None
            self._snapshot_best()
            self.best_epoch = 0
            self.best_time = 0
            self.best_loss_rel = torch.ones_like(self.best_loss)
            self.best_func_count = 2 * batch_size * self.n_flow
            self.history = []
            del XJ, X, J
        if run is not None and log:
            run.log_scalar("training.int_loss", self.best_loss.tolist(), 0)
        self.int_loss = self.best_loss

        state = EpochState(self.int_loss, preburn_time, kill_counter, impr_ratio)
        params = [p for p in self.model.parameters() if p.requires_grad]
        i = epoch_start - 1
        # Small minibatches leave most of the GPU idle (a 2000-point minibatch of the README example occupies 16 CTAs), and the
        # minibatches of an epoch are independent (fresh points, own BatchNorm statistics, one backward of their mean): they
        # run side by side on a few streams - forwards and, since autograd replays a node on the stream of its forward, the
        # backwards too.  BatchNorm's running statistics are a sequential recurrence, so every concurrent forward writes
        # its update into a private zeroed buffer (FlowSpec.bn_override: m * batch statistic) and the buffers are folded
        # back in minibatch order afterwards: rs <- (1-m)^n rs + sum_k (1-m)^(n-1-k) priv_k - what n sequential forwards
        # produce.  NIS_TRAIN_STREAMS=0 (or ``minibatch_streams = 0`` on the manager) keeps the minibatches in sequence.
        n_par = int(os.environ.get("NIS_TRAIN_STREAMS", getattr(self, "minibatch_streams", 8)))
        par_streams = None
        fold_weights = {}                              # minibatch count -> (1-m)^(n-1-k), made outside any graph capture
        if n_par > 1 and hasattr(self.model, "spec") and len(my_minibatches) > 1 and mini_batch_size <= 16384 and dev.type == "cuda":
            par_streams = [torch.cuda.Stream(device=dev) for _ in range(min(n_par, len(my_minibatches)))]

        def one_minibatch(preburn, w):
            XJ = self.model(self.format_input(w, dev))
            X = XJ[:, :-1].detach()                 # the sample is fixed, the Jacobian is optimised
            if preburn:
                fres = f(w)
                fXJ = torch.mul(fres, XJ[:, -1]) / maxf
                integ_k = torch.mean(fres) / n_minibatches
                err_k = torch.var(fres) / n_minibatches
            else:
                fres = torch.mul(f(X), XJ[:, -1])
                fXJ = fres / maxf
                integ_k = torch.mean(fres.detach()) / n_minibatches
                err_k = torch.var(fres.detach()) / n_minibatches
            loss_k = torch.var(fXJ) if loss_mode == "var" else torch.mean((fXJ * maxf) ** 2)
            var_k = torch.var(fXJ.detach() ** 2) / mini_batch_size
            return loss_k, var_k, integ_k, err_k

        def epoch_body(preburn, batches):
            """One epoch up to and including backward (manager.py:219-278): fresh uniform minibatches, the variance of
            f J / maxf per minibatch, one backward of their mean.  Returns (loss, var, integ, err) tensors."""
            ws = [self._uniform(mini_batch_size, dev, generator=gen) for _ in batches]     # drawn in minibatch order
            parts = []
            if par_streams is not None and len(ws) > 1:
                spec = self.model.spec()
                cur = torch.cuda.current_stream(dev)
                bn_flat = spec.bn_arena.get(dev)
                priv = torch.zeros(len(ws), bn_flat.numel(), dtype=bn_flat.dtype, device=dev)
                used = par_streams[:min(len(par_streams), len(ws))]
                for st_ in used:
                    st_.wait_stream(cur)
                try:
                    for k, w in enumerate(ws):
                        with torch.cuda.stream(used[k % len(used)]):
                            spec.bn_override = priv[k]
                            parts.append(one_minibatch(preburn, w))
                finally:
                    spec.bn_override = None
                for st_ in used:
                    cur.wait_stream(st_)
                # fold the private running-statistics updates back, in minibatch order
                keep = 1.0 - float(spec.desc.bn_momentum)
                wts = fold_weights.get(len(ws))
                if wts is None:                        # (first reached in an eager warm-up epoch, never under capture)
                    wts = torch.tensor([keep ** (len(ws) - 1 - k) for k in range(len(ws))], dtype=bn_flat.dtype, device=dev)
                    fold_weights[len(ws)] = wts
                with torch.no_grad():
                    bn_flat.mul_(keep ** len(ws)).add_(torch.matmul(wts, priv))
                    spec.bn_arena.write_back(bn_flat)
                    nbt = spec.nbt_arena.get(dev)
                    nbt += len(ws)
                    spec.nbt_arena.write_back(nbt)
            else:
                for w in ws:
                    parts.append(one_minibatch(preburn, w))
            loss, var, integ_e, err_e = 0, 0, 0, 0
            for loss_k, var_k, integ_k, err_k in parts:
                loss = loss + loss_k
                var = var + var_k
                integ_e = integ_e + integ_k
                err_e = err_e + err_k
            if torch.is_tensor(loss):
                loss = loss / n_minibatches
                loss.backward()
            return loss, var, integ_e, err_e

        # The epoch body is latency-bound for small minibatches (README example: ~40 launches per minibatch, more
        # host time than device time).  After two eager epochs in the current mode it is captured into a CUDA graph
        # and replayed (the optimizer step and the bookkeeping stay eager).  A user integrand that cannot be
        # captured (host synchronisation, Python-side state) makes the capture fail, and training continues eagerly;
        # ``cuda_graph_epochs = False`` on the manager (or NIS_TRAIN_GRAPH=0) turns it off.
        graphs = {}                                    # preburner flag -> (CUDAGraph, static outputs)
        eager_epochs = {True: 0, False: 0}
        side = None                                    # warm-up epochs and capture share one side stream (so that the
                                                       # parameters' AccumulateGrad nodes live on the capture stream)
        use_graph = world == 1 and getattr(self, "cuda_graph_epochs", True) and \
            os.environ.get("NIS_TRAIN_GRAPH", "1") != "0" and len(my_minibatches) > 0

        for i in epoch_progress:
            mode = bool(state.preburner)
            if use_graph and mode not in graphs and eager_epochs[mode] >= 2:
                optimizer_object.zero_grad(set_to_none=True)
                try:
                    torch.cuda.synchronize(dev)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=side):
                        outs = epoch_body(mode, my_minibatches)
                    graphs[mode] = (g, outs, [p.grad for p in params])
                except Exception as exc:               # not capturable: stay eager for good
                    use_graph = False
                    graphs.clear()
                    optimizer_object.zero_grad(set_to_none=True)
                    torch.cuda.synchronize(dev)
                    warnings.warn("the training epoch could not be captured into a CUDA graph (%s); continuing "
                                  "eagerly" % (str(exc).splitlines()[0] if str(exc) else type(exc).__name__))
            if use_graph and mode in graphs:
                g, outs, grads = graphs[mode]
                for p, gr in zip(params, grads):       # the graph writes its gradients into these very tensors
                    p.grad = gr
                g.replay()
                loss, var, integ_e, err_e = outs
                loss = loss.clone()
            elif use_graph:
                if side is None:
                    side = torch.cuda.Stream(device=dev)
                optimizer_object.zero_grad()
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    loss, var, integ_e, err_e = epoch_body(mode, minibatch_progress)
                torch.cuda.current_stream(dev).wait_stream(side)
                eager_epochs[mode] += 1
            else:
                optimizer_object.zero_grad()
                loss, var, integ_e, err_e = epoch_body(mode, minibatch_progress)
            if torch.is_tensor(integ_e):
                integ[i + 1] += integ_e
                err[i + 1] += err_e
            if not torch.is_tensor(loss):              # a rank without minibatches this epoch
                loss = sum(p.sum() for p in params) * 0.0
                var = torch.zeros((), device=dev, dtype=torch.double)
                loss.backward()
            if world > 1:
                self._allreduce_grads(params)
                stats = torch.stack((loss.detach().double(), var.double(), integ[i + 1].double(), err[i + 1].double()))
                dist.all_reduce(stats)
                loss = stats[0].to(loss.dtype)
                var = stats[1]
                integ[i + 1], err[i + 1] = stats[2], stats[3]
            optimizer_object.step()
            loss = loss.detach()
            var = float(var)

            self.history.append(loss) if hasattr(self, "history") else None
            if pretty_progressbar and rank == 0:
                epoch_progress.set_description("Loss: {0:.3e} | Epoch".format(loss))
            if run is not None and log:
                run.log_scalar("training.loss", loss.tolist(), i)
                run.log_scalar("training.loss_rel", (loss / self.int_loss).tolist(), i)

            if save_best or log:
                self.best_func_count = self.best_func_count + batch_size
            if state.improved(loss, save_best or log):
                self.best_loss = loss
                self.best_var = var
                self.best_loss_rel = loss / self.int_loss
                self._snapshot_best()
                self.best_epoch = i
                self.best_time = (datetime.datetime.utcnow() - run.start_time).total_seconds() if run is not None else 0
            if state.advance(i, loss):
                break

        # ---- optional tail integration with the best model in eval mode (manager.py:332-350) -------
        endpoint = i + 1
        with torch.no_grad():
            if integrate and endpoint < epochs - 1:
                model = self.best_model.eval()
                for s in range(endpoint, epochs):
                    for t in my_minibatches:
                        w = self._uniform(mini_batch_size, dev, generator=gen)
                        XJ = model(self.format_input(w, dev)).detach()
                        fres = torch.mul(f(XJ[:, :-1]), XJ[:, -1])
                        integ[s + 1] += torch.mean(fres) / (n_minibatches * math.sqrt(mini_batch_size))
                        err[s + 1] += torch.std(fres) / n_minibatches
                    self.best_func_count = self.best_func_count + batch_size
                if world > 1:
                    tail = torch.stack((integ[endpoint + 1:], err[endpoint + 1:]))
                    dist.all_reduce(tail)
                    integ[endpoint + 1:], err[endpoint + 1:] = tail[0], tail[1]
        self.integ_tot = torch.sum(integ / err) / torch.sum(1 / err)
        self.err_tot = torch.sqrt(1 / torch.sum(1 / err))

        if run is not None and integrate:
            run.log_scalar("training.integ", self.integ_tot.tolist(), 0)
            run.log_scalar("training.err", self.err_tot.tolist(), 0)
        if log and rank == 0:
            try:
                torch.save({'best_epoch': self.best_epoch, 'best_loss': self.best_loss, 'int_loss': self.int_loss,
                            'best_loss_rel': self.best_loss_rel, 'best_func_count': self.best_func_count,
                            'model_state_dict': self.best_model.state_dict(), 'integ': self.integ_tot,
                            'err': self.err_tot}, filename)
            except Exception:
                print("Torch save not possible")
        if integrate:
            return (self.integ_tot.detach().tolist(), self.err_tot.detach().tolist())
        return (0, 0)

    @staticmethod
    def _allreduce_grads(params):
        """One sum-allreduce of all gradients (NCCL over NVLink when launched one rank per GPU).

        ``nis_flow_backward`` writes the whole parameter gradient into ONE flat float32 buffer and autograd hands
        the parameters views of it, so after ``zero_grad()`` (set_to_none) + ``backward()`` every ``p.grad`` is a
        slice of that buffer: it is all-reduced in place with a single collective and no host-side fan-out.  Any
        other situation (gradients accumulated elsewhere, a user calling ``zero_grad(set_to_none=False)``) takes
        the gather / scatter path."""
        flat = getattr(params[0].grad, "_base", None) if params and params[0].grad is not None else None
        if flat is not None and flat.dim() == 1 and flat.is_contiguous():
            lo, hi, isz, covered = flat.data_ptr(), flat.data_ptr() + flat.numel() * flat.element_size(), \
                flat.element_size(), 0
            for p in params:
                g = p.grad
                if g is None or g.dtype != flat.dtype or not g.is_contiguous() or getattr(g, "_base", None) is not flat \
                        or not (lo <= g.data_ptr() and g.data_ptr() + g.numel() * isz <= hi):
                    flat = None
                    break
                covered += g.numel()
            if flat is not None and covered == flat.numel():
                dist.all_reduce(flat)
                return
        for p in params:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        grads = [p.grad for p in params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat)
        torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])

    def _sync_model(self):
        """Ranks must start from identical weights: broadcast parameters and BN buffers from rank 0."""
        tensors = list(self.model.parameters()) + [b for b in self.model.buffers() if b.dtype.is_floating_point]
        flat = torch.cat([t.detach().reshape(-1).float() for t in tensors])
        dist.broadcast(flat, 0)
        off = 0
        with torch.no_grad():
            for t in tensors:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n

    # ------------------------------------------------------------------------------------------
    def _weight_moments(self, f, nitn, neval, dev):
        """[nitn, 6] float64 on the host: (sum, sum of squares, n, max, min, number of non-finite) of f(x) J(x) over
        neval points per iteration through ``best_model``.  The points of an iteration are split over the ranks
        (disjoint slices of one Philox stream); one sum- and one max-allreduce combine them."""
        dev = _device(0 if dev is None else dev)
        rank, world = _world()
        first, share = shard_bounds(neval, rank, world)
        lib = _cabi.lib()
        moments = torch.zeros(nitn, 6, dtype=torch.double, device=dev)
        rws = torch.empty(lib.nis_reduce_workspace_bytes(), dtype=torch.uint8, device=dev)
        w = torch.empty(share, self.n_flow, device=dev)                     # float32 like manager.py:390
        seed = int(torch.empty((), dtype=torch.int64).random_().item())     # drawn from torch's CPU generator
        if world > 1:
            s = torch.tensor([seed], device=dev)
            dist.broadcast(s, 0)
            seed = int(s.item())
        with torch.no_grad(), torch.cuda.device(dev):
            for i in range(nitn):
                _cabi.check(lib.nis_uniform_fill(_cabi.ptr(w), _cabi.F32, w.numel(), seed,
                                                 (i * neval + first) * self.n_flow, _cabi.stream_ptr(dev)),
                            "nis_uniform_fill")
                X = self.best_model(self.format_input(w, dev)).detach()
                fres = (f(X[:, :-1]) * X[:, -1]).contiguous()
                _cabi.check(lib.nis_reduce_stats(_cabi.ptr(fres), _cabi.dtype_code(fres), fres.numel(),
                                                 _cabi.ptr(moments[i]), 0, _cabi.ptr(rws), rws.numel(),
                                                 _cabi.stream_ptr(dev)), "nis_reduce_stats")
        if world > 1:
            sums = moments[:, [0, 1, 2, 5]].contiguous()
            ext = torch.stack((moments[:, 3], -moments[:, 4]), 1)
            dist.all_reduce(sums)
            dist.all_reduce(ext, op=dist.ReduceOp.MAX)
            moments = torch.stack((sums[:, 0], sums[:, 1], sums[:, 2], ext[:, 0], -ext[:, 1], sums[:, 3]), 1)
        moments = moments.cpu()
        self.n_nonfinite = int(moments[:, 5].sum())
        if self.n_nonfinite:
            warnings.warn("%d of %d integrand weights f(x)*J(x) are inf / NaN: the estimate below is not finite "
                          "(the reference returns the same NaN silently)" % (self.n_nonfinite, int(moments[:, 2].sum())),
                          NonFiniteWeightWarning, stacklevel=3)
        return moments

    def integrate(self, f, nitn, neval, dev=None):
        """nitn independent estimates of neval points through ``best_model``, combined by inverse
        variance (manager.py:380-405, including its error formula).  With torch.distributed the neval
        points of every iteration are split over the ranks and (sum, sum of squares, n) are allreduced.
        Non-finite weights are counted (``self.n_nonfinite``) and reported by a ``NonFiniteWeightWarning``."""
        if self.best_model is None:
            print("No model has been trained")
            return (0, 0)
        neval, nitn = int(neval), int(nitn)
        moments = self._weight_moments(f, nitn, neval, dev)
        n = moments[:, 2]
        mean = moments[:, 0] / n
        var = (moments[:, 1] - n * mean ** 2) / (n - 1)                      # unbiased, like torch.var
        mean, var = mean.float(), var.float()                                # manager.py:391-392 holds float32
        sig = torch.sum(mean / var) / torch.sum(1 / var)
        sig_err = torch.sqrt(1 / torch.sum(1 / var)) / np.sqrt(neval * nitn)
        return (sig, sig_err)

    def weight_statistics(self, f, neval, dev=None):
        """What the reference's experiment harness computes after training from one batch of ``neval`` points
        through ``best_model`` (utils/experiment_mg.py:66-76,101): final variance ``v_var`` (unbiased), ``w_max``,
        ``w_mean`` and the unweighting efficiency ``w_mean / w_max`` — as ONE fused reduction on the device
        (``nis_reduce_stats``), sharded over the ranks like ``integrate``."""
        if self.best_model is None:
            print("No model has been trained")
            return None
        m = self._weight_moments(f, 1, int(neval), dev)[0]
        n = float(m[2])
        mean = float(m[0]) / n
        var = (float(m[1]) - n * mean * mean) / (n - 1)
        return {"v_var": var, "w_mean": mean, "w_max": float(m[3]), "w_min": float(m[4]),
                "unweighting_efficiency": mean / float(m[3]) if float(m[3]) != 0 else float("nan"),
                "n": int(n), "n_nonfinite": int(m[5])}


def _finish_model(manager, model, dev):
    manager._model = model
    if torch.cuda.is_available():
        model.to(_device(dev))
    manager.best_model = manager.model
    if torch.cuda.is_available():                        # one pass forward, like the reference factories
        w = torch.rand(5, manager.n_flow, dtype=torch.double)
        with torch.no_grad():
            model(manager.format_input(w, _device(dev)))


class PWLinManager(BasicManager):
    """Piecewise-linear coupling cells with cyclic roll layers (manager.py:456-499).

    Hyperparameters: n_pass_through, n_cells, n_bins, NN (hidden widths), roll_step.
    """

    def create_model(self, n_pass_through, n_cells, n_bins, NN, roll_step):
        model = FlowSequential(self.n_flow)
        for i_cell in range(n_cells):
            model.add_module(str(i_cell), PWLin(flow_size=self.n_flow, pass_through_size=n_pass_through,
                                                n_bins=n_bins, NN_layers=NN))
            # The reference registers every roll under the one name "roll" (manager.py:492): add_module
            # replaces it in place, so a single roll survives, right after cell 0.  Reproduced as is.
            model.add_module("roll", RollLayer(roll_step))
        _finish_model(self, model, 0)


class AffineManager(BasicManager):
    """Affine coupling cells with cyclic roll layers (manager.py:411-453).  Same topology as PWLinManager -
    the reference registers every roll under the one name "roll", so a single roll survives right after cell 0.

    Hyperparameters: n_pass_through, n_cells, NN (hidden widths), roll_step.  (The reference's create_model raises torch's
    mixed-dtype error in its trial pass on current torch, after ``_model`` is set - the same quirk as PWLinManager's; this
    one runs the trial pass.)"""

    def create_model(self, n_pass_through, n_cells, NN, roll_step):
        model = FlowSequential(self.n_flow)
        for i_cell in range(n_cells):
            model.add_module(str(i_cell), AffineCoupling(flow_size=self.n_flow, pass_through_size=n_pass_through,
                                                         NN_layers=NN))
            model.add_module("roll", RollLayer(roll_step))
        _finish_model(self, model, 0)


class PWQuadManager(BasicManager):
    """Piecewise-quadratic coupling cells; roll layout for n_flow <= 7, binary-mask layout above
    (manager.py:502-600).  Hyperparameters: n_cells, n_bins, NN (hidden widths)."""

    def create_model(self, n_cells, n_bins, NN, dev=0):
        d = self.n_flow
        if n_cells < 2 * np.ceil(np.log2(d)) and n_cells < d:            # manager.py:526-534
            n_cells = d if d <= 6 else (6 if d == 7 else int(2 * np.ceil(np.log2(d))))
            print("Adjusted # coupling cells to " + str(n_cells))
        model = FlowSequential(d)

        def roll_cells(first, count, n_pass_through):
            # `count` cells, each followed by a unit roll; the last roll restores the original order
            for k in range(count):
                name = str(first + k)
                model.add_module(name, PWQuad(flow_size=d, pass_through_size=n_pass_through, n_bins=n_bins,
                                              NN_layers=NN))
                shift = 1 if k < count - 1 else d - ((count - 1) % d)
                model.add_module("roll" + name, RollLayer(shift))

        if d <= 7:
            roll_cells(0, n_cells, 1 if d <= 6 else 2)
        else:
            nbits = len(get_bin(d - 1, 0))
            dims_bin = torch.IntTensor([get_bin(i, nbits) for i in range(d)])
            for c in range(2 * nbits):
                masker = MaskLayer(dims_bin, c, "cpu")
                model.add_module("mask" + str(c), masker)
                model.add_module(str(c), PWQuad(flow_size=d, pass_through_size=masker.pass_through,
                                                n_bins=n_bins, NN_layers=NN))
                model.add_module("demask" + str(c), DeMaskLayer(masker.feeder, masker.trafoer))
            roll_cells(2 * nbits, n_cells - 2 * nbits, int(d / 2))
        _finish_model(self, model, dev)
        return
